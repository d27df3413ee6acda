"""Fused kernel-matrix assembly: achieved GB/s (algorithmic bytes as stored) by kernel expression and mode."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import portfoliooptgp_b200 as gpflow
from portfoliooptgp_b200 import ops
from portfoliooptgp_b200.kernels import compile_kernel

N, D = 8192, 8
X, _ = bench.make_c2(n=N, d=D)
Xd = torch.as_tensor(X, device="cuda")
Xs = torch.as_tensor(bench.make_c2(seed=9, n=16384, d=D)[0], device="cuda")
K = gpflow.kernels
kernels = {
    "SE": K.SquaredExponential(),
    "Matern52": K.Matern52(),
    "SE+Matern52+Linear": K.SquaredExponential() + K.Matern52() + K.Linear(),
    "Exp[0:7]*Exp[7]": K.Exponential(active_dims=slice(0, 7)) * K.Exponential(active_dims=slice(7, 8)),
    "Exp+Periodic(SE)[7]+Linear": K.Exponential() + K.Periodic(K.SquaredExponential(active_dims=[7])) + K.Linear(),
}
eng = ops.shared_engine(0); ops.sync_stream(eng)
out = torch.empty((N, 16384), dtype=torch.float64, device="cuda")
peak = bench.measured_peaks()["hbm_gbs"]
res = {}
for name, k in kernels.items():
    ck = compile_kernel(k, D); eng.set_kernel(ck.spec); th = ck.theta()
    for mode, label, n2, bytes_ in ((1, "lower", N, 8.0 * N * (N + 1) / 2), (2, "symmetric_full", N, 8.0 * N * N), (0, "cross_8192x16384", 16384, 8.0 * N * 16384)):
        def run():
            eng.assemble(th, Xd.data_ptr(), N, Xs.data_ptr() if mode == 0 else None, n2, D, out.data_ptr(), out.shape[1], mode, 1e-2 if mode else 0.0)
        for _ in range(2): run()
        torch.cuda.synchronize(); best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        gbs = (bytes_ + 8.0 * D * (N + n2)) / (best * 1e-3) / 1e9
        res[f"{name}|{label}"] = {"ms": best, "GBps": gbs, "frac_of_hbm_copy": gbs / peak}
        print(f"{name:28s} {label:18s} {best:7.3f} ms  {gbs:7.1f} GB/s  {gbs / peak:5.2f} of HBM", flush=True)
json.dump(res, open("gpurun_out/assembly_bench.json", "w"), indent=1)
