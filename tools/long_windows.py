"""BatchedGPR on windows LONGER than 128 rows (gpb_gpr_lml_grad_many): LML+gradient evaluations per second by
window length and by the number of handles working side by side; then one full ragged expanding-window fit."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import portfoliooptgp_b200 as gpflow
from tests.helpers import make_multi_input

out = {"evals_per_s": {}}
D, B = 8, 64
K = gpflow.kernels
for N in (160, 256, 512, 1000):
    X, Y = make_multi_input(3, N + B - 1, D)
    Xb = np.stack([X[i:i + N] for i in range(B)]); Yb = np.stack([Y[i:i + N, 0] for i in range(B)])
    row = {}
    for nh in (1, 2, 4, 8, 16):
        os.environ["GPB_MANY_HANDLES"] = str(nh)
        k = K.Exponential(active_dims=slice(0, D - 1), lengthscales=1.3) * K.Exponential(active_dims=slice(D - 1, D), variance=0.8)
        m = gpflow.BatchedGPR(Xb, Yb, k, noise_variance=1e-2)
        m.lml_and_grads()
        reps = 3
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            m.lml_and_grads()
        dt = (time.perf_counter() - t0) / reps
        row[str(nh)] = round(B / dt, 1)
        del m
    out["evals_per_s"][str(N)] = row
    print(N, row, flush=True)
# the reference's loop shape: expanding windows, one more row per test day (Multi-Input_GPR/main.py:414-456)
os.environ["GPB_MANY_HANDLES"] = "8"
first, count = 200, 32
Xf, Yf = make_multi_input(5, first + count, D)
nrows = first + np.arange(count)
Xfull = np.repeat(Xf[None, :nrows.max()], count, axis=0); Yfull = np.repeat(Yf[None, :nrows.max(), 0], count, axis=0)
k = K.Exponential(active_dims=slice(0, D - 1), lengthscales=1.3) * K.Exponential(active_dims=slice(D - 1, D), variance=0.8)
m = gpflow.BatchedGPR(Xfull, Yfull, k, noise_variance=1e-3, nrows=nrows, train_noise=False)
t0 = time.perf_counter(); res = m.fit(maxiter=100); t_batch = time.perf_counter() - t0
nfev = int(sum(r.nfev for r in res))
# the same fits one after the other, as the reference's loop does
t0 = time.perf_counter()
nfev_seq = 0
for b, i in enumerate(nrows):
    kb = K.Exponential(active_dims=slice(0, D - 1), lengthscales=1.3) * K.Exponential(active_dims=slice(D - 1, D), variance=0.8)
    g = gpflow.models.GPR((Xf[:i], Yf[:i]), kernel=kb, noise_variance=1e-3)
    gpflow.set_trainable(g.likelihood, False)
    r = gpflow.optimizers.Scipy().minimize(g.training_loss, g.trainable_variables, options=dict(maxiter=100))
    nfev_seq += int(r.nfev)
    assert r.nit == res[b].nit, (b, r.nit, res[b].nit)
t_seq = time.perf_counter() - t0
out["expanding_windows_200_to_231_rows_32_gps"] = {"lock_step_8_handles_s": round(t_batch, 4), "one_after_the_other_s": round(t_seq, 4),
                                                   "nfev": nfev, "nfev_sequential": nfev_seq, "same_iteration_counts": True}
print(json.dumps(out))
