"""A/B of the look-ahead chain at the bottom of the factorisation (gpb_set_option(h, 5, x)): LML + gradient and
value-only at several N, potrf at 8192, C1-size evaluation latency."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import portfoliooptgp_b200 as gpflow

out = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(f, reps):
    for _ in range(3):
        r = f()
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        r = f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


for n in [int(a) for a in sys.argv[1:]] or [1000, 2048, 4096, 8192]:
    d = 1 if n == 1000 else 8
    if n == 1000:
        from tests.helpers import make_c1
        X, Y = make_c1(n)
        k = gpflow.kernels.SquaredExponential() + gpflow.kernels.Periodic(gpflow.kernels.SquaredExponential())
    else:
        X, Y = bench.make_c2(n=n)
        k = gpflow.kernels.SquaredExponential() + gpflow.kernels.Matern52() + gpflow.kernels.Linear()
    m = gpflow.models.GPR((X, Y), kernel=k, noise_variance=1e-2)
    eng = m._get_engine()
    row = {}
    for mode in (0, 1):
        eng.set_option(eng.OPTION_CHAIN, mode)
        tg, rg = timed(m.lml_and_constrained_grads, 20 if n <= 4096 else 10)
        tv, rv = timed(lambda: float(m.log_marginal_likelihood()), 20 if n <= 4096 else 10)
        row["chain" if mode else "recursion"] = {"lml_grad_ms": tg, "lml_value_ms": tv, "lml": rg[0]}
    eng.set_option(eng.OPTION_CHAIN, 1)
    out[n] = row
print(json.dumps(out))
