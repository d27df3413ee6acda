import os, sys
os.environ["GPB_BATCHED_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import portfoliooptgp_b200 as gpflow
X, Y = bench.make_c2(seed=3, n=128 + 300, d=8)
B = 296
Xb = np.stack([X[i:i + 128] for i in range(B)]); Yb = np.stack([Y[i:i + 128, 0] for i in range(B)])
K = gpflow.kernels
for name, k in {"exp*exp": K.Exponential(active_dims=slice(0, 7)) * K.Exponential(active_dims=slice(7, 8)),
                "se+m52+lin": K.SquaredExponential() + K.Matern52() + K.Linear(), "se": K.SquaredExponential()}.items():
    m = gpflow.BatchedGPR(Xb, Yb, k, noise_variance=1e-2)
    print(name, file=sys.stderr)
    for _ in range(2):
        m.lml_and_grads()
