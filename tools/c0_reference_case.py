"""The reference's own small case (N = 89 daily points, the 8 candidate kernels of
GPR/main.py:105-114, sigma^2 = 1e-5 frozen, L-BFGS-B maxiter = 100, in-sample predict_f, MSE select)
through the drop-in API, next to the CPU oracle doing the same fits with SciPy."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import portfoliooptgp_b200 as gpflow
rng = np.random.default_rng(7)
N = 89
X = ((np.arange(N) - (N - 1) / 2) / np.std(np.arange(N)))[:, None]
Y = np.cumsum(rng.normal(size=(N, 1)) * 0.01, axis=0); Y = (Y - Y.mean()) / Y.std()
def kernels():
    K = gpflow.kernels
    return [K.SquaredExponential(), K.Matern12(), K.RationalQuadratic(), K.Exponential(), K.SquaredExponential() + K.Matern12(),
            K.Exponential() + K.Periodic(K.SquaredExponential()) + K.Linear(), K.Exponential() + K.Periodic(K.SquaredExponential()),
            K.SquaredExponential() * K.Matern12()]
def train_all():
    best = (None, np.inf)
    for k in kernels():
        m = gpflow.models.GPR(data=(X, Y), kernel=k)
        m.likelihood.variance.assign(1e-5)
        gpflow.set_trainable(m.likelihood.variance, False)
        try:
            gpflow.optimizers.Scipy().minimize(m.training_loss, m.trainable_variables, options=dict(maxiter=100))
            mean, _ = m.predict_f(X)
            mse = float(np.mean((mean.numpy() - Y) ** 2))
        except gpflow.CholeskyError:
            mse = np.inf
        if mse < best[1]:
            best = (type(k).__name__, mse)
    return best
train_all()
t0 = time.perf_counter(); best = train_all(); dt = time.perf_counter() - t0
print(f"reference small case: 8 kernels x (fit maxiter=100 + predict) at N={N}: {dt*1e3:.1f} ms wall, best {best}")
