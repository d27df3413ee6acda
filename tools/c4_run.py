"""BASELINE config C4: large exact GP, N = 65536, D = 4 (3 return columns + time), kernel SE + Matern52,
sigma^2 = 1e-2, fixed theta: blocked fp64 Cholesky (factor only for the value / cold predict_f, factor +
inverse for the gradient) and predict_f at 16384 held-out points
on ONE GPU.  Prints timings and size-independent sanity properties (no CPU oracle at this size)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import portfoliooptgp_b200 as gpflow

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
Ns = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
D = 4
X, Y = bench.make_c2(seed=4, n=N + Ns, d=D)
perm = np.random.default_rng(4).permutation(N + Ns)
Xtr, Ytr, Xte, Yte = X[perm[:N]], Y[perm[:N]], X[perm[N:]], Y[perm[N:]]
k = gpflow.kernels.SquaredExponential(lengthscales=1.5) + gpflow.kernels.Matern52(variance=0.5, lengthscales=3.0)
m = gpflow.models.GPR((Xtr, Ytr), kernel=k, noise_variance=1e-2)
gpflow.models.set_output_device("cuda")
out = {"N": N, "Ns": Ns, "D": D}
torch.cuda.synchronize(); t0 = time.perf_counter()
lml = float(m.log_marginal_likelihood())
torch.cuda.synchronize(); out["lml_s"] = time.perf_counter() - t0; out["lml"] = lml
t0 = time.perf_counter()
mean, var = m.predict_f(Xte)      # same theta as the LML evaluation above: the factorisation is reused
torch.cuda.synchronize(); out["predict_f_after_lml_s"] = time.perf_counter() - t0
m._fact = None                    # forget it: the stand-alone cost of predict_f (factorisation included)
t0 = time.perf_counter()
mean2, var2 = m.predict_f(Xte)
torch.cuda.synchronize(); out["predict_f_s"] = time.perf_counter() - t0
out["predict_reuse_identical"] = bool(torch.equal(mean, mean2) and torch.equal(var, var2))
t0 = time.perf_counter()
lml2, g, gn = m.lml_and_constrained_grads()
torch.cuda.synchronize(); out["lml_grad_s"] = time.perf_counter() - t0
out["lml_repeat_rel_diff"] = abs(lml2 - lml) / abs(lml)
out["factor_only_tflops"] = (float(N) ** 3 / 3) / out["lml_s"] / 1e12   # value only: the factor alone (csrc/cholesky.cu factor_L)
out["lml_grad_tflops"] = float(N) ** 3 / out["lml_grad_s"] / 1e12
mv, vv = mean.cpu().numpy(), var.cpu().numpy()
out["var_min"], out["var_max"] = float(vv.min()), float(vv.max())
out["test_rmse"] = float(np.sqrt(np.mean((mv - Yte) ** 2)))
out["test_rmse_of_zero_predictor"] = float(np.sqrt(np.mean(Yte ** 2)))
_, tv = m.predict_f(Xtr[:2048])
out["train_var_below_noise"] = bool((tv.cpu().numpy() < 1e-2).all() and (tv.cpu().numpy() > 0).all())
out["mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
out["grad"] = [float(v) for v in g] + [float(gn)]
print(json.dumps(out))
