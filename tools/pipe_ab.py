"""A/B of the pipelined two-partition factorisation against the single-partition recursion (C2 workload)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import portfoliooptgp_b200 as gpflow

out = {}
for n in [int(a) for a in sys.argv[1:]] or [4096, 8192, 16384]:
    X, Y = bench.make_c2(n=n)
    k = gpflow.kernels.SquaredExponential() + gpflow.kernels.Matern52() + gpflow.kernels.Linear()
    m = gpflow.models.GPR((X, Y), kernel=k, noise_variance=1e-2)
    eng = m._get_engine()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    row = {}
    for mode in (0, 1):
        eng.set_option(eng.OPTION_PIPELINE, mode)
        for _ in range(3):
            r = m.lml_and_constrained_grads()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            r = m.lml_and_constrained_grads()
        e1.record(); torch.cuda.synchronize()
        row["pipeline" if mode else "recursion"] = {"ms": e0.elapsed_time(e1) / 10, "lml": r[0]}
    eng.set_option(eng.OPTION_PIPELINE, 0)
    out[n] = row
print(json.dumps(out))
