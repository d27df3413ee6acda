import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import portfoliooptgp_b200 as gpflow
from portfoliooptgp_b200 import ops
from portfoliooptgp_b200.kernels import compile_kernel
N, D = 8192, 8
X, _ = bench.make_c2(n=N, d=D)
Xd = torch.as_tensor(X, device="cuda")
k = gpflow.kernels.SquaredExponential() if len(sys.argv) < 2 else gpflow.kernels.SquaredExponential() + gpflow.kernels.Matern52() + gpflow.kernels.Linear()
ck = compile_kernel(k, D); eng = ops.shared_engine(0); ops.sync_stream(eng); eng.set_kernel(ck.spec)
out = torch.empty((N, N), dtype=torch.float64, device="cuda")
for _ in range(3):
    eng.assemble(ck.theta(), Xd.data_ptr(), N, None, N, D, out.data_ptr(), N, 1, 1e-2)
torch.cuda.synchronize()
print(float(out[5, 3]))
