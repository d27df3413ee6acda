"""A/B: programmatic dependent launch on/off for the C2 evaluation and a C1-size evaluation."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import portfoliooptgp_b200 as gpflow
for n, d, reps in ((8192, 8, 10), (1000, 8, 200)):
    X, Y = bench.make_c2(n=n, d=d)
    k = gpflow.kernels.SquaredExponential() + gpflow.kernels.Matern52() + gpflow.kernels.Linear()
    m = gpflow.models.GPR((X, Y), kernel=k, noise_variance=1e-2)
    eng = m._get_engine()
    for pdl in (1, 0, 1, 0):
        eng.set_option(eng.OPTION_PDL, pdl)
        for _ in range(3): out = m.lml_and_constrained_grads()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): out = m.lml_and_constrained_grads()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
        print(f"N={n} pdl={pdl}: {dt*1e3:.3f} ms/eval  lml={out[0]:.10f}", flush=True)
