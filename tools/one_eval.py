"""One warm-up + EVALS LML+grad evaluations of the C2 workload (for ncu launch lists / captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import portfoliooptgp_b200 as gpflow

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
evals = int(sys.argv[2]) if len(sys.argv) > 2 else 1
X, Y = bench.make_c2(n=n)
k = gpflow.kernels.SquaredExponential() + gpflow.kernels.Matern52() + gpflow.kernels.Linear()
m = gpflow.models.GPR((X, Y), kernel=k, noise_variance=1e-2)
for _ in range(1 + evals):
    out = m.lml_and_constrained_grads()
print(out[0], m._engine.launch_count())
