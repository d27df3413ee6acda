"""Full lock-step fit of configuration C3 (5120 GPs: 20 assets x 64 windows x 4 noise restarts, N = 128,
D = 8, Exponential * Exponential, trainable noise): wall time, rounds, where the time goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import portfoliooptgp_b200 as gpflow
B = int(sys.argv[1]) if len(sys.argv) > 1 else 5120
X, Y = bench.make_c2(seed=3, n=128 + 64 * 20 + 8, d=8)
nwin = B // 4
Xw = np.stack([X[i:i + 128] for i in range(nwin)]); Yw = np.stack([Y[i:i + 128, 0] for i in range(nwin)])
Xb = np.repeat(Xw, 4, axis=0); Yb = np.repeat(Yw, 4, axis=0)
noise0 = np.tile(np.array([1e-5, 1e-3, 1e-1, 1.0]), nwin)
K = gpflow.kernels
k = K.Exponential(active_dims=slice(0, 7)) * K.Exponential(active_dims=slice(7, 8))
m = gpflow.BatchedGPR(Xb, Yb, k, noise_variance=noise0, train_noise=True)
dev_t = [0.0]; calls = [0]
orig = m.loss_and_grads_unconstrained
def timed(U, idx):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = orig(U, idx)
    torch.cuda.synchronize(); dev_t[0] += time.perf_counter() - t0; calls[0] += 1
    return out
m.loss_and_grads_unconstrained = timed
workers = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if workers > 1:   # start the pool outside the timed region (interpreter + numpy/scipy import per worker)
    from portfoliooptgp_b200 import _lbfgsb_pool
    _lbfgsb_pool.get_workers(workers)
t0 = time.perf_counter()
res = m.fit(maxiter=100, workers=workers)
dt = time.perf_counter() - t0
nit = np.array([r.nit for r in res]); ok = np.array([r.success for r in res])
print(f"C3 full fit (workers={workers}): {B} GPs, {dt:.2f} s wall, {calls[0]} lock-step rounds, device+copies {dev_t[0]:.2f} s, host (SciPy setulb) {dt - dev_t[0]:.2f} s; "
      f"nit mean {nit.mean():.1f} max {nit.max()}, converged {ok.mean()*100:.1f} %, {B/dt:.0f} fits/s")
