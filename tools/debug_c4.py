"""Step-by-step check of the N = 65 536 comparison path (tests/test_gpu_baseline_sizes.py::test_c4_n65536_vs_cusolver)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import portfoliooptgp_b200 as gp
from portfoliooptgp_b200 import ops
from tests.helpers import make_multi_input

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
D, Ns, noise = 4, 512, 1e-2
X, Y = make_multi_input(4, N + Ns, D)
Xs, X, Y = X[N:], X[:N], Y[:N]
K = gp.kernels
k = K.SquaredExponential(variance=1.0, lengthscales=1.5) + K.Matern52(variance=0.5, lengthscales=2.5)

def step(name, f):
    t = time.time()
    out = f()
    torch.cuda.synchronize()
    print(name, "ok %.2fs" % (time.time() - t), flush=True)
    return out

Kfull = step("assemble mode2", lambda: ops.kernel_matrix(k, X, diag_add=noise))
print("  sym check", float((Kfull[:100, 60000 % N:60000 % N + 50] - Kfull[60000 % N:60000 % N + 50, :100].T).abs().max()) if N > 100 else "")
Ks = step("assemble cross", lambda: ops.kernel_matrix(k, X, Xs))
kss = step("kdiag", lambda: ops.kernel_diag(k, Xs))
L = step("torch cholesky", lambda: torch.linalg.cholesky(Kfull))
