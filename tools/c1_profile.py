"""cProfile of the C1 fit (N = 1000, D = 1, SE + Periodic(SE), L-BFGS-B): where the host time per evaluation goes."""
import os, sys, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import portfoliooptgp_b200 as gpflow
rng = np.random.default_rng(1)
X = ((np.arange(1000.0) - 499.5) / 288.8)[:, None]
Y = np.sin(3 * X) + 0.3 * rng.normal(size=(1000, 1)); Y = (Y - Y.mean()) / Y.std()
def fit():
    k = gpflow.kernels.SquaredExponential() + gpflow.kernels.Periodic(gpflow.kernels.SquaredExponential())
    m = gpflow.models.GPR((X, Y), kernel=k, noise_variance=1e-2)
    res = gpflow.optimizers.Scipy().minimize(m.training_loss, m.trainable_variables, options=dict(maxiter=100))
    return res
fit()
t0 = time.perf_counter(); res = fit(); dt = time.perf_counter() - t0
print(f"fit: {dt*1e3:.1f} ms, nfev {res.nfev}, {dt/res.nfev*1e3:.3f} ms per evaluation")
pr = cProfile.Profile(); pr.enable(); res = fit(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
pstats.Stats(pr).sort_stats("cumtime").print_stats(28)
