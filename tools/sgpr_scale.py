"""SGPR at scale: N = 262144, M = 1024, D = 8, SquaredExponential -- ELBO + full gradient timing and the
algorithmic FP64 rate of its four M x M x N products."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import portfoliooptgp_b200 as gpflow
N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
X, Y = bench.make_c2(seed=6, n=N, d=8)
Z = X[np.random.default_rng(6).choice(N, M, replace=False)].copy()
m = gpflow.models.SGPR((X, Y), kernel=gpflow.kernels.SquaredExponential(), inducing_variable=Z, noise_variance=0.1)
clo = m.training_loss_closure()
tv = m.trainable_variables
for _ in range(2):
    loss, grads = clo.value_and_grads(tv)
torch.cuda.synchronize(); t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    loss, grads = clo.value_and_grads(tv)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
flops = 2.0 * M * M * N * (0.5 + 0.5 + 1.0 + 0.5)   # V (tri A), V V^T (lower), G1 V, Kuf_bar V^T (lower)
print(f"SGPR N={N} M={M}: loss {loss:.6f}, {dt*1e3:.2f} ms per ELBO+grad, {flops/dt/1e12:.1f} TFLOP/s algorithmic on the M^2 N products")
t0 = time.perf_counter(); e = float(m.elbo()); torch.cuda.synchronize(); print(f"ELBO only: {(time.perf_counter()-t0)*1e3:.2f} ms, elbo {e:.6f}")
Xs = X[:20000]
t0 = time.perf_counter(); fm, fv = m.predict_f(Xs); torch.cuda.synchronize(); print(f"predict_f at 20000 points: {(time.perf_counter()-t0)*1e3:.2f} ms, var range {float(fv.min()):.3e} .. {float(fv.max()):.3e}")
