"""One warm-up + EVALS LML+grad evaluations of the C1 workload (N = 1000, D = 1, SE + Periodic(SE)) for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import portfoliooptgp_b200 as gpflow

evals = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rng = np.random.default_rng(1)
X = ((np.arange(1000.0) - 499.5) / 288.8)[:, None]
Y = np.sin(3 * X) + 0.3 * rng.normal(size=(1000, 1)); Y = (Y - Y.mean()) / Y.std()
k = gpflow.kernels.SquaredExponential() + gpflow.kernels.Periodic(gpflow.kernels.SquaredExponential())
m = gpflow.models.GPR((X, Y), kernel=k, noise_variance=1e-2)
for _ in range(1 + evals):
    out = m.lml_and_constrained_grads()
print(out[0], m._engine.launch_count())
