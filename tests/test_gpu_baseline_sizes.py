"""GPU parity AT THE BASELINE SIZES (BASELINE.json configs C1, C2, C4, C5; VERDICT r01 missing #1-#4),
each against the CPU oracle on the same seeded inputs, at the north-star tolerances (1e-9 relative on
LML / ELBO / predictive mean and variance, 1e-7 on gradients, optimiser end points within 1e-6 after the
same iterations) -- or, where fp64 itself cannot meet them, at the bound the extended-precision truth
gives for fp64 LAPACK on that case (tests/test_truth.py, C1 at sigma^2 = 1e-5: predictive mean 5e-8).

Sizes were chosen so that each oracle call finishes in seconds on the GPU box's host cores; the largest
(C4 at N = 65 536) has no affordable CPU oracle and is compared with torch.linalg.cholesky (cuSOLVER) as
an independent comparison point, never a product path."""
import math

import numpy as np
import pytest
import scipy.optimize

from oracle import gpflow_oracle as O
from oracle import gpflow_oracle_torch as T
from tests.helpers import make_c1, make_multi_input, record_parity, to_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _direct_form():
    O.set_distance_form("direct")   # the distance form the CUDA kernels evaluate
    yield
    O.set_distance_form("gram")


def _c1_kernel(gp):
    K = gp.kernels
    return K.SquaredExponential() + K.Periodic(K.SquaredExponential())


def _c1_oracle():
    return O.Sum([O.Leaf("se"), O.Periodic(O.Leaf("se"), 1.0)])


# ---- C1: N = 1000, D = 1, SE + Periodic(SE), sigma^2 frozen (GPR/model_trainer.py:15-20) -------------------
@pytest.mark.parametrize("tag,s2,mean_bar", [("1e-2", 1e-2, 1e-9), ("1e-5", 1e-5, 5e-8)])
def test_c1_lml_grad_predict_match_oracle(gp, tag, s2, mean_bar):
    X, Y = make_c1(1000)
    Xs = np.concatenate([X[::7], X[-1:] + (X[1] - X[0]) * np.arange(1, 31)[:, None]])   # in-sample + 30 future days
    m = gp.models.GPR(data=(X, Y), kernel=_c1_kernel(gp))
    m.likelihood.variance.assign(s2)
    gp.set_trainable(m.likelihood.variance, False)
    lml, g, gn = m.lml_and_constrained_grads()
    mean, var = m.predict_f(Xs, full_cov=False)
    ko = _c1_oracle()
    l0, g0, n0 = O.gpr_lml_and_grad(ko, X, Y, s2)
    m0, v0 = O.gpr_predict_f(ko, X, Y, s2, Xs)
    gscale = max(np.max(np.abs(g0)), abs(n0))
    err = {"lml_rel": abs(lml - l0) / abs(l0), "grad_rel_to_max": float(max(np.max(np.abs(g - g0)), abs(gn - n0)) / gscale),
           "mean_rel_to_max": float(np.max(np.abs(mean.numpy() - m0)) / np.max(np.abs(m0))),
           "var_abs": float(np.max(np.abs(var.numpy() - v0))), "var_max": float(np.max(v0))}
    record_parity("c1_n1000_vs_oracle|" + tag, err)
    assert err["lml_rel"] <= 1e-9, err
    assert err["grad_rel_to_max"] <= 1e-7, err
    assert err["mean_rel_to_max"] <= mean_bar, err
    assert err["var_abs"] <= 1e-9 * max(1.0, err["var_max"]), err


def _c1_fit(gp, X, Y, s2, maxiter):
    m = gp.models.GPR(data=(X, Y), kernel=_c1_kernel(gp))
    m.likelihood.variance.assign(s2)
    gp.set_trainable(m.likelihood.variance, False)
    variables = m.trainable_variables
    x0 = gp.optimizers.Scipy.initial_parameters(variables)
    res = gp.optimizers.Scipy().minimize(m.training_loss, variables, options=dict(maxiter=maxiter))
    return m, x0, res


def _c1_oracle_objective(X, Y, s2):
    ko = _c1_oracle()

    def fun(u):
        O.set_theta(ko, O.softplus(u))
        l, g, _ = O.gpr_lml_and_grad(ko, X, Y, s2)
        return -l, -g * O.sigmoid(u)

    return ko, fun


@pytest.mark.parametrize("tag,s2", [("1e-2", 1e-2), ("1e-5", 1e-5)])
def test_c1_scipy_iterates_match_oracle_lbfgs(gp, tag, s2):
    """GPR/model_trainer.py:15-19 at C1's size, same SciPy L-BFGS-B, same x0 and variable order, CUDA objective
    vs oracle objective: the iterates agree to 1e-6 "after the same iterations" (north_star) for as long as the
    comparison is well-posed -- the first 3 iterations (10 evaluations).  This objective (a flat valley: the
    SE variance runs to 0) amplifies a perturbation of the iterate ~100-1000x per iteration: the ORACLE ITSELF,
    Gram-form vs direct distances (two fp64 LAPACK runs, f differing by 3e-14 relative), is 5e-10 / 4e-7 apart
    after 3 iterations and 4e-7 / 2.5e-3 after 4 (sigma^2 = 1e-2 / 1e-5), and ends the maxiter = 100 run after
    71 vs 56 iterations 0.9 apart in x; f scaled by (1 + 1e-14) or another BLAS do the same (measured,
    DESIGN.md section 10)."""
    X, Y = make_c1(1000)
    m, x0, res = _c1_fit(gp, X, Y, s2, maxiter=3)
    ko, fun = _c1_oracle_objective(X, Y, s2)
    ref = scipy.optimize.minimize(fun, x0, jac=True, method="L-BFGS-B", options=dict(maxiter=3))
    rec = {"nit_gpu": int(res.nit), "nit_oracle": int(ref.nit), "nfev_gpu": int(res.nfev), "nfev_oracle": int(ref.nfev),
           "x_max_abs_diff": float(np.max(np.abs(res.x - ref.x))), "fun_rel_diff": float(abs(res.fun - ref.fun) / abs(ref.fun))}
    record_parity("c1_n1000_lbfgs_3_iterations|" + tag, rec)
    assert res.nit == ref.nit and res.nfev == ref.nfev, rec
    assert rec["x_max_abs_diff"] < 1e-6, rec
    assert rec["fun_rel_diff"] < (1e-9 if s2 >= 1e-2 else 5e-8), rec


@pytest.mark.parametrize("tag,s2", [("1e-2", 1e-2), ("1e-5", 1e-5)])
def test_c1_full_fit_end_point_is_the_oracles_optimum_too(gp, tag, s2):
    """The full maxiter = 100 fit + in-sample predict_f (GPR/model_trainer.py:19-20).  The end points of two fp64
    runs are not comparable coordinate by coordinate (see above), so the CUDA end point is checked where it
    stands: the oracle's objective, gradient and prediction AT that point agree with the CUDA ones, the loss
    there is as low as the oracle's own run reaches (within 0.5 %), and the fit converged."""
    X, Y = make_c1(1000)
    m, x0, res = _c1_fit(gp, X, Y, s2, maxiter=100)
    ko, fun = _c1_oracle_objective(X, Y, s2)
    f_at, g_at = fun(res.x)
    ref = scipy.optimize.minimize(fun, x0, jac=True, method="L-BFGS-B", options=dict(maxiter=100))
    loss, grads = m.training_loss_closure().value_and_grads(m.trainable_variables)
    g_gpu = np.concatenate([np.atleast_1d(g) for g in grads])
    mean, var = m.predict_f(X)
    O.set_theta(ko, O.softplus(res.x))
    m0, v0 = O.gpr_predict_f(ko, X, Y, s2, X)
    rec = {"nit_gpu": int(res.nit), "nit_oracle": int(ref.nit), "fun_gpu": float(res.fun), "fun_oracle_own_run": float(ref.fun),
           "oracle_f_at_gpu_x_rel": float(abs(f_at - loss) / abs(f_at)),
           "oracle_g_at_gpu_x_rel_to_max": float(np.max(np.abs(g_at - g_gpu)) / max(1.0, np.max(np.abs(g_at)))),
           "mean_rel_to_max": float(np.max(np.abs(mean.numpy() - m0)) / np.max(np.abs(m0))),
           "var_abs": float(np.max(np.abs(var.numpy() - v0)))}
    record_parity("c1_n1000_full_fit|" + tag, rec)
    assert res.success, res.message
    assert rec["oracle_f_at_gpu_x_rel"] <= 1e-9 and rec["oracle_g_at_gpu_x_rel_to_max"] <= 1e-7, rec
    assert res.fun <= ref.fun + 5e-3 * abs(ref.fun), rec
    assert rec["mean_rel_to_max"] <= (1e-9 if s2 >= 1e-2 else 5e-8) and rec["var_abs"] <= 1e-9, rec


# ---- C2: N = 8192, D = 8, SE + Matern52 + Linear (north-star sum kernel), sigma^2 = 1e-2 ---------------------
def test_c2_full_size_lml_grad_match_oracle(gp):
    N, D, noise = 8192, 8, 1e-2
    X, Y = make_multi_input(2, N, D)
    K = gp.kernels
    k = K.SquaredExponential(lengthscales=1.3) + K.Matern52(variance=0.7, lengthscales=2.0) + K.Linear(variance=0.2)
    m = gp.models.GPR((X, Y), kernel=k, noise_variance=noise)
    lml, g, gn = m.lml_and_constrained_grads()
    l0, g0, n0 = O.gpr_lml_and_grad(to_oracle(k), X, Y, noise)
    gscale = max(1.0, np.max(np.abs(g0)), abs(n0))
    rec = {"lml_rel": abs(lml - l0) / abs(l0), "grad_rel_to_max": float(max(np.max(np.abs(g - g0)), abs(gn - n0)) / gscale)}
    record_parity("c2_n8192_vs_oracle", rec)
    assert rec["lml_rel"] <= 1e-9, rec
    assert rec["grad_rel_to_max"] <= 1e-7, rec
    # the LML-only entry point (no K^-1) gives the same value
    assert abs(float(m.log_marginal_likelihood()) - lml) <= 1e-12 * abs(lml)


# ---- C4: large exact GP, D = 4, SE + Matern52, sigma^2 = 1e-2, fixed theta (Multi-Input_GPR/main.py:421-434) ---
def _c4_kernel(gp):
    K = gp.kernels
    return K.SquaredExponential(variance=1.0, lengthscales=1.5) + K.Matern52(variance=0.5, lengthscales=2.5)


def test_c4_n16384_factor_and_predict_match_oracle(gp):
    N, D, Ns, noise = 16384, 4, 2048, 1e-2
    X, Y = make_multi_input(4, N + Ns, D)
    Xs, X, Y = X[N:], X[:N], Y[:N]
    k = _c4_kernel(gp)
    m = gp.models.GPR((X, Y), kernel=k, noise_variance=noise)
    mean, var = m.predict_f(Xs)                        # cold start: factorisation + solves
    lml = float(m.log_marginal_likelihood())
    ko = to_oracle(k)
    l0 = O.gpr_lml(ko, X, Y, noise)
    m0, v0 = O.gpr_predict_f(ko, X, Y, noise, Xs)
    rec = {"lml_rel": abs(lml - l0) / abs(l0), "mean_rel_to_max": float(np.max(np.abs(mean.numpy() - m0)) / np.max(np.abs(m0))),
           "var_rel": float(np.max(np.abs(var.numpy() - v0) / np.abs(v0)))}
    record_parity("c4_n16384_vs_oracle", rec)
    assert rec["lml_rel"] <= 1e-9 and rec["mean_rel_to_max"] <= 1e-9 and rec["var_rel"] <= 1e-9, rec
    # predict after an objective evaluation at the same parameters reuses the factorisation: same numbers
    mean2, var2 = m.predict_f(Xs)
    assert np.max(np.abs(mean2.numpy() - mean.numpy())) <= 1e-12 * np.max(np.abs(m0))
    assert np.max(np.abs(var2.numpy() - var.numpy())) <= 1e-12


def test_c4_n65536_vs_cusolver(gp):
    """BASELINE config 4 at its full size.  No CPU oracle is affordable (a 34 GB matrix, 9.4e13 flop);
    the comparison point is cuSOLVER through torch.linalg.cholesky on the engine-assembled matrix
    (independent factorisation and solves; the assembly itself is oracle-checked at N = 16 384 above)."""
    import torch
    from portfoliooptgp_b200 import ops
    free, total = torch.cuda.mem_get_info()
    if free < 150e9:
        pytest.skip("needs ~140 GB of free HBM")
    N, D, Ns, noise = 65536, 4, 512, 1e-2
    X, Y = make_multi_input(4, N + Ns, D)
    Xs, X, Y = X[N:], X[:N], Y[:N]
    k = _c4_kernel(gp)
    Kfull = ops.kernel_matrix(k, X, diag_add=noise)               # [N, N] symmetric, 34 GB
    Ks = ops.kernel_matrix(k, X, Xs)                               # [N, Ns]
    kss = ops.kernel_diag(k, Xs)
    # torch.linalg.cholesky on the whole 34 GB matrix faults inside the library at this size (an illegal
    # address, reproduced by tools/debug_c4.py on the assembled matrix alone), so the comparison factorisation
    # is a right-looking block Cholesky over 16 384-wide panels made of library pieces: cuSOLVER potrf on the
    # diagonal blocks, cuBLAS triangular solves and matmul for the rest -- none of this repo's kernels.
    nb = 16384
    Yd = torch.as_tensor(Y, device="cuda")
    B = torch.cat([Yd, Ks.contiguous()], dim=1)                     # right-hand sides [N, 1 + Ns]
    logdet = 0.0
    for o in range(0, N, nb):
        Lkk = torch.linalg.cholesky(Kfull[o:o + nb, o:o + nb])
        logdet += float(torch.log(torch.diagonal(Lkk)).sum())
        B[o:o + nb] = torch.linalg.solve_triangular(Lkk, B[o:o + nb], upper=False)
        if o + nb < N:
            L21 = torch.linalg.solve_triangular(Lkk, Kfull[o + nb:, o:o + nb].T, upper=False).T.contiguous()   # [rest, nb]
            B[o + nb:] -= L21 @ B[o:o + nb]
            for c in range(o + nb, N, nb):                          # trailing update, lower block columns
                Kfull[c:, c:c + nb] -= L21[c - o - nb:] @ L21[c - o - nb:c - o].T
            del L21
        del Lkk
    a, A = B[:, :1], B[:, 1:]
    ref_lml = float(-0.5 * (a * a).sum() - 0.5 * N * math.log(2 * math.pi) - logdet)
    ref_mean = (A.T @ a).cpu().numpy()
    ref_var = (kss - (A * A).sum(0)).cpu().numpy()[:, None]
    del Kfull, B, Ks, a, A
    torch.cuda.empty_cache()
    m = gp.models.GPR((X, Y), kernel=k, noise_variance=noise)
    lml = float(m.log_marginal_likelihood())
    mean, var = m.predict_f(Xs)
    rec = {"lml_rel": abs(lml - ref_lml) / abs(ref_lml),
           "mean_rel_to_max": float(np.max(np.abs(mean.numpy() - ref_mean)) / np.max(np.abs(ref_mean))),
           "var_rel": float(np.max(np.abs(var.numpy() - ref_var) / np.abs(ref_var)))}
    record_parity("c4_n65536_vs_cusolver", rec)
    del m
    torch.cuda.empty_cache()
    assert rec["lml_rel"] <= 1e-9 and rec["mean_rel_to_max"] <= 1e-9 and rec["var_rel"] <= 1e-9, rec


# ---- C5: SVGP M = 2048, D = 8 (test_scripts/SVGP.py:515-533 call pattern at the BASELINE size) -------------------
def _c5_setup(gp, B, seed=5):
    M, D = 2048, 8
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((B, D))
    w = rng.standard_normal((D, 1))
    Y = np.sin(X @ w) + 0.1 * rng.standard_normal((B, 1))
    Z = rng.standard_normal((M, D))
    qmu = 0.3 * rng.standard_normal((M, 1))
    qs = (0.7 * np.eye(M) + 0.01 * np.tril(rng.standard_normal((M, M))))[None]
    k = gp.kernels.SquaredExponential(lengthscales=2.0)
    m = gp.models.SVGP(kernel=k, likelihood=gp.likelihoods.Gaussian(variance=1e-2), inducing_variable=Z,
                       num_data=16 * 2 ** 20, q_mu=qmu, q_sqrt=qs)
    return m, k, X, Y, Z, qmu, qs


def test_c5_elbo_at_full_minibatch_matches_oracle(gp):
    B = 65536
    m, k, X, Y, Z, qmu, qs = _c5_setup(gp, B)
    elbo = float(m.elbo((X, Y)))
    e0 = O.svgp_elbo(to_oracle(k), Z, qmu, qs, 1e-2, X, Y, num_data=16 * 2 ** 20)
    rec = {"elbo_gpu": elbo, "elbo_oracle": float(e0), "elbo_rel": abs(elbo - e0) / abs(e0)}
    record_parity("c5_m2048_b65536_elbo_vs_oracle", rec)
    assert rec["elbo_rel"] <= 1e-9, rec


def test_c5_gradients_at_m2048_match_torch_oracle(gp):
    B, M = 8192, 2048
    m, k, X, Y, Z, qmu, qs = _c5_setup(gp, B, seed=6)
    ko = to_oracle(k)
    e0, g0 = T.svgp_elbo_and_grad(ko, Z, qmu, qs, 1e-2, X, Y, num_data=16 * 2 ** 20)
    variables = m.trainable_variables
    loss, grads = m.training_loss_closure((X, Y)).value_and_grads(variables)
    by_var = {id(v): g for v, g in zip(variables, grads)}
    rec = {"elbo_rel": abs(-loss - e0) / abs(e0)}
    got_theta = np.concatenate([np.atleast_1d(-by_var[id(p.unconstrained_variable)]
                                              / p.transform.forward_grad(p.unconstrained_variable.numpy())) for p in k.parameters])
    rec["theta"] = float(np.max(np.abs(got_theta - g0["theta"])) / max(1.0, np.max(np.abs(g0["theta"]))))
    gZ = -by_var[id(m.inducing_variable.Z.unconstrained_variable)]
    rec["Z"] = float(np.max(np.abs(gZ - g0["Z"])) / max(1.0, np.max(np.abs(g0["Z"]))))
    gq = -by_var[id(m.q_mu.unconstrained_variable)]
    rec["q_mu"] = float(np.max(np.abs(gq - g0["q_mu"])) / max(1.0, np.max(np.abs(g0["q_mu"]))))
    gs = -by_var[id(m.q_sqrt.unconstrained_variable)]
    want = g0["q_sqrt"][:, np.tril_indices(M)[0], np.tril_indices(M)[1]]
    rec["q_sqrt"] = float(np.max(np.abs(gs - want)) / max(1.0, np.max(np.abs(want))))
    pv = m.likelihood.variance
    gn = -by_var[id(pv.unconstrained_variable)] / pv.transform.forward_grad(pv.unconstrained_variable.numpy())
    rec["noise"] = float(abs(float(gn) - g0["noise"]) / max(1.0, abs(g0["noise"])))
    record_parity("c5_m2048_b8192_grads_vs_torch_oracle", rec)
    assert rec["elbo_rel"] <= 1e-9, rec
    for name in ("theta", "Z", "q_mu", "q_sqrt", "noise"):
        assert rec[name] <= 1e-7, rec
