"""gpb_potrf (factor only, N^3/3 flop) against cuSOLVER's potrf (torch.linalg.cholesky) on the same matrix,
and the value-only / value+gradient objective at the same N.  Comparison points only."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import portfoliooptgp_b200 as gpflow
from portfoliooptgp_b200 import ops

out = {}
for n in [int(a) for a in sys.argv[1:]] or [4096, 8192, 16384]:
    X, Y = bench.make_c2(n=n)
    k = gpflow.kernels.SquaredExponential() + gpflow.kernels.Matern52() + gpflow.kernels.Linear()
    K0 = ops.kernel_matrix(k, X, diag_add=1e-2).contiguous()
    eng = ops.shared_engine(0)
    A = torch.empty_like(K0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(f, reps=5):
        best = 1e9
        for _ in range(reps):
            A.copy_(K0); torch.cuda.synchronize()
            e0.record(); f(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    ops.sync_stream(eng)
    t_gpb = timed(lambda: eng.potrf(A.data_ptr(), n, n))
    Lg = torch.tril(A).clone()
    t_cus = timed(lambda: torch.linalg.cholesky(A, out=A) if False else torch.linalg.cholesky_ex(A, check_errors=False))
    Lc = torch.linalg.cholesky(K0)
    m = gpflow.models.GPR((X, Y), kernel=k, noise_variance=1e-2)
    m.lml_and_constrained_grads(); float(m.log_marginal_likelihood())

    def ev(f, reps=5):
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize(); e0.record(); f(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    t_val = ev(lambda: m.log_marginal_likelihood())
    t_grad = ev(lambda: m.lml_and_constrained_grads())
    out[n] = {"gpb_potrf_ms": t_gpb, "cusolver_potrf_ms": t_cus, "potrf_tflops": n ** 3 / 3 / t_gpb / 1e9,
              "max_abs_L_diff": float((Lg - Lc).abs().max()), "lml_value_only_ms": t_val, "lml_grad_ms": t_grad}
    del K0, A, Lg, Lc, m
    torch.cuda.empty_cache()
print(json.dumps(out))
