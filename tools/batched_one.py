"""Two batched LML+grad launches (C3 shape: N=128, D=8, Exponential*Exponential) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import portfoliooptgp_b200 as gpflow
X, Y = bench.make_c2(seed=3, n=128 + 600, d=8)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
Xb = np.stack([X[i:i + 128] for i in range(B)]); Yb = np.stack([Y[i:i + 128, 0] for i in range(B)])
K = gpflow.kernels
k = K.Exponential(active_dims=slice(0, 7)) * K.Exponential(active_dims=slice(7, 8))
m = gpflow.BatchedGPR(Xb, Yb, k, noise_variance=1e-2)
for _ in range(2):
    f = m.lml_and_grads()
torch.cuda.synchronize()
print("lml[0..2] =", f[0][:3], "info max", f[3].max())
