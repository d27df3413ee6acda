"""GPU parity: GPR LML, gradient, predict_f / predict_y and L-BFGS end points vs the CPU oracle.

Tolerances (BASELINE.json north_star): relative 1e-9 on LML, predictive mean and variance; 1e-7
on gradients; optimiser end points within 1e-6 after the same iterations.  The oracle is run in
the direct-difference distance form (the form the CUDA kernels evaluate); headline configs use
sigma^2 >= 1e-2 so that cond(K) <~ 1e6 and the bar is meaningful."""
import numpy as np
import pytest
import scipy.optimize

from oracle import gpflow_oracle as O
from tests.helpers import kernel_zoo, make_multi_input, to_oracle

pytestmark = pytest.mark.gpu

RTOL_VALUE = 1e-9
RTOL_GRAD = 1e-7


@pytest.fixture(autouse=True)
def _direct_form():
    O.set_distance_form("direct")
    yield
    O.set_distance_form("gram")


def _rel(a, b, floor=1e-12):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))


@pytest.mark.parametrize("D,N", [(1, 89), (1, 300), (4, 200), (8, 513)])
def test_lml_and_grad_match_oracle(gp, D, N):
    X, Y = make_multi_input(21, N, D)
    noise = 1e-2
    for name, k in kernel_zoo(D).items():
        m = gp.models.GPR((X, Y), kernel=k, noise_variance=noise)
        lml, g_theta, g_noise = m.lml_and_constrained_grads()
        ko = to_oracle(k)
        l0, g0, n0 = O.gpr_lml_and_grad(ko, X, Y, noise)
        assert abs(lml - l0) <= RTOL_VALUE * abs(l0), (name, lml, l0)
        assert float(m.log_marginal_likelihood()) == pytest.approx(lml, rel=1e-13)
        gscale = max(1.0, np.max(np.abs(g0)))
        assert np.max(np.abs(g_theta - g0)) <= RTOL_GRAD * gscale, (name, g_theta, g0)
        assert abs(g_noise - n0) <= RTOL_GRAD * max(1.0, abs(n0)), name


@pytest.mark.parametrize("D,N,Ns", [(1, 120, 57), (8, 400, 130)])
def test_predict_matches_oracle(gp, D, N, Ns):
    X, Y = make_multi_input(31, N, D)
    Xs, _ = make_multi_input(32, Ns, D)
    noise = 1e-2
    for name, k in kernel_zoo(D).items():
        m = gp.models.GPR((X, Y), kernel=k, noise_variance=noise)
        mean, var = m.predict_f(Xs, full_cov=False)
        ymean, yvar = m.predict_y(Xs)
        m0, v0 = O.gpr_predict_f(to_oracle(k), X, Y, noise, Xs)
        scale_m = np.max(np.abs(m0))
        assert mean.shape == (Ns, 1) and var.shape == (Ns, 1)
        assert np.max(np.abs(mean.numpy() - m0)) <= RTOL_VALUE * scale_m, name
        assert _rel(var.numpy(), v0, floor=1e-3) <= RTOL_VALUE * 10, name
        assert np.allclose(yvar.numpy() - var.numpy(), noise, rtol=0, atol=1e-15)
        assert np.array_equal(ymean.numpy(), mean.numpy())


def test_scipy_endpoint_matches_oracle_lbfgs(gp):
    """Same SciPy L-BFGS-B, same x0 ordering, oracle objective vs GPU objective: end points agree
    (reference call pattern GPR/model_trainer.py:15-19 with maxiter=100)."""
    X, Y = make_multi_input(41, 150, 1)
    k = gp.kernels.SquaredExponential() + gp.kernels.Matern12()
    m = gp.models.GPR(data=(X, Y), kernel=k)
    m.likelihood.variance.assign(1e-2)
    gp.set_trainable(m.likelihood.variance, False)
    variables = m.trainable_variables
    x0 = gp.optimizers.Scipy.initial_parameters(variables)
    res = gp.optimizers.Scipy().minimize(m.training_loss, variables, options=dict(maxiter=100))

    ko = to_oracle(gp.kernels.SquaredExponential() + gp.kernels.Matern12())

    def fun(u):
        theta = O.softplus(u)
        O.set_theta(ko, theta)
        l, g, _ = O.gpr_lml_and_grad(ko, X, Y, 1e-2)
        return -l, -g * O.sigmoid(u)

    ref = scipy.optimize.minimize(fun, x0, jac=True, method="L-BFGS-B", options=dict(maxiter=100))
    assert res.nit == ref.nit
    assert np.max(np.abs(res.x - ref.x)) < 1e-6
    assert abs(res.fun - ref.fun) < 1e-6 * max(1.0, abs(ref.fun))
    # parameters were assigned back
    assert np.allclose(gp.optimizers.Scipy.initial_parameters(variables), res.x)


def test_reference_call_pattern_model_trainer(gp):
    """GPR/model_trainer.py:10-26 verbatim apart from the import: shared kernel instances
    warm-start across fits (SURVEY.md section 3 aliasing note)."""
    gpflow = gp
    X, Y = make_multi_input(51, 89, 1)
    kernels = [gpflow.kernels.SquaredExponential(), gpflow.kernels.Exponential() + gpflow.kernels.Linear()]
    best = None
    for kernel in kernels:
        model = gpflow.models.GPR(data=(X, Y), kernel=kernel)
        model.likelihood.variance.assign(1e-5)
        gpflow.set_trainable(model.likelihood.variance, False)
        opt = gpflow.optimizers.Scipy()
        opt.minimize(model.training_loss, model.trainable_variables, options=dict(maxiter=100))
        mean_test, _ = model.predict_f(X)
        mse = float(np.mean((Y - mean_test.numpy()) ** 2))
        if best is None or mse < best[0]:
            best = (mse, kernel, model)
    assert best[0] < 1.0
    assert float(kernels[0].lengthscales.numpy()) != 1.0  # trained in place


def test_non_pd_raises(gp):
    X = np.zeros((50, 1)); Y = np.zeros((50, 1))
    m = gp.models.GPR((X, Y), kernel=gp.kernels.Linear(), noise_variance=2e-6)
    m.likelihood.variance.unconstrained_variable.assign(-800.0)  # variance -> lower bound 1e-6 exactly
    X2 = np.ones((50, 1)) * 1e8
    m2 = gp.models.GPR((X2, Y), kernel=gp.kernels.Linear(), noise_variance=2e-6)
    with pytest.raises(gp.CholeskyError):
        float(m2.log_marginal_likelihood())


def test_mean_functions_train_and_predict(gp):
    """Constant / Linear / Polynomial(2) mean functions (test_scripts/GPFlow.py:186-190, GPR.py:103):
    LML and gradients w.r.t. the mean parameters against the oracle (mean enters only through Y - m(X))."""
    X, Y = make_multi_input(61, 150, 1)
    Y = Y + 0.7 * X + 0.3                     # give the mean something to explain
    noise = 1e-2
    k = gp.kernels.SquaredExponential(lengthscales=0.7)
    ko = to_oracle(k)
    for mf, m_of_x, dparams in [
        (gp.mean_functions.Constant(0.2), lambda X: 0.2 + 0 * X, lambda a, X: [a.sum()]),
        (gp.mean_functions.Linear(np.array([[0.5]]), np.array([0.1])), lambda X: 0.5 * X + 0.1,
         lambda a, X: [(a * X[:, 0]).sum(), a.sum()]),
        (gp.mean_functions.Polynomial(2, w=[0.1, 0.4, -0.05]), lambda X: 0.1 + 0.4 * X - 0.05 * X ** 2,
         lambda a, X: [a.sum(), (a * X[:, 0]).sum(), (a * X[:, 0] ** 2).sum()]),
    ]:
        m = gp.models.GPR((X, Y), kernel=k, mean_function=mf, noise_variance=noise)
        resid = Y - m_of_x(X)
        l0 = O.gpr_lml(ko, X, resid, noise)
        assert abs(float(m.log_marginal_likelihood()) - l0) <= 1e-9 * abs(l0)
        L = O.gpr_cholesky(ko, X, noise)
        import scipy.linalg as sla
        alpha = sla.cho_solve((L, True), resid)[:, 0]
        variables = m.trainable_variables
        loss, grads = m.training_loss_closure().value_and_grads(variables)
        by_var = {id(v): g for v, g in zip(variables, grads)}
        got = np.concatenate([-by_var[id(p.unconstrained_variable)].reshape(-1) for p in mf.parameters])
        want = np.array(dparams(alpha, X), dtype=np.float64)
        assert np.max(np.abs(got - want)) <= 1e-7 * max(1.0, np.max(np.abs(want)))
        # predict_f adds the mean back
        Xs = np.linspace(-1, 1, 7)[:, None]
        mean, _ = m.predict_f(Xs)
        m0, _ = O.gpr_predict_f(ko, X, resid, noise, Xs)
        assert np.max(np.abs(mean.numpy() - (m0 + m_of_x(Xs)))) <= 1e-9 * max(1.0, np.max(np.abs(m0)))
    # and the optimiser can train them
    mf = gp.mean_functions.Linear()
    m = gp.models.GPR((X, Y), kernel=gp.kernels.SquaredExponential(), mean_function=mf, noise_variance=noise)
    before = float(m.training_loss())
    res = gp.optimizers.Scipy().minimize(m.training_loss, m.trainable_variables, options=dict(maxiter=30))
    assert res.fun < before


def test_two_devices_in_one_process(gp):
    """One process, engines on two GPUs (per-device kernel attributes, workspaces, streams): same answers."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    X, Y = make_multi_input(3, 700, 8)
    k = gp.kernels.SquaredExponential() + gp.kernels.Matern52() + gp.kernels.Linear()
    out = []
    for dev in (0, 1):
        m = gp.models.GPR((X, Y), kernel=k, noise_variance=1e-2, device=dev)
        lml, g, gn = m.lml_and_constrained_grads()
        fm, fv = m.predict_f(X[:40])
        bg = gp.BatchedGPR(np.stack([X[i:i + 128] for i in range(4)]), np.stack([Y[i:i + 128, 0] for i in range(4)]), k,
                           noise_variance=0.1, device=dev)
        out.append((lml, g, gn, np.asarray(fm), np.asarray(fv), bg.lml_and_grads()[0]))
        # the C-ABI restores the caller's current device on return (a call on a cuda:1 model must
        # not move torch's current device)
        assert torch.cuda.current_device() == 0, dev
    assert out[0][0] == out[1][0] and out[0][2] == out[1][2]
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][3], out[1][3]) and np.array_equal(out[0][4], out[1][4])
    assert np.array_equal(out[0][5], out[1][5])


def test_predict_reuses_the_factorisation_only_when_it_may(gp):
    """predict_y after predict_f (GPR/predictor.py:6-7), and predict after an objective evaluation at the
    same theta, skip the factorisation (fewer launches, identical numbers); another model in between, a
    parameter change or new data force the full path."""
    X, Y = make_multi_input(8, 900, 3)
    Xs = np.random.default_rng(2).normal(size=(60, 3))
    k = gp.kernels.SquaredExponential(lengthscales=0.9) + gp.kernels.Matern12(variance=0.4)
    m = gp.models.GPR((X, Y), kernel=k, noise_variance=0.05)
    eng = m._get_engine()

    def launches(fn):
        n0 = eng.launch_count()
        out = fn()
        return out, eng.launch_count() - n0

    (f1, v1), n_full = launches(lambda: m.predict_f(Xs))
    (f2, v2), n_reuse = launches(lambda: m.predict_f(Xs))
    assert n_reuse < n_full - 5, (n_full, n_reuse)
    assert np.array_equal(np.asarray(f1), np.asarray(f2)) and np.array_equal(np.asarray(v1), np.asarray(v2))
    (ym, yv), n_y = launches(lambda: m.predict_y(Xs))
    assert n_y == n_reuse
    np.testing.assert_allclose(np.asarray(yv), np.asarray(v1) + 0.05, rtol=1e-15)
    # objective evaluation, then predict at the same theta: reuse
    m.training_loss_closure().value_and_grads(m.trainable_variables)
    (f3, v3), n3 = launches(lambda: m.predict_f(Xs))
    assert n3 == n_reuse and np.array_equal(np.asarray(f3), np.asarray(f1)) and np.array_equal(np.asarray(v3), np.asarray(v1))
    # another model on the same engine in between: full path, still the right answer
    other = gp.models.GPR((X[:500], Y[:500]), kernel=gp.kernels.Matern32(), noise_variance=0.1)
    other.predict_f(Xs)
    (f4, v4), n4 = launches(lambda: m.predict_f(Xs))
    assert n4 == n_full and np.array_equal(np.asarray(f4), np.asarray(f1))
    # parameter change: full path, different answer, matches a fresh model
    k.kernels[0].lengthscales.assign(1.7)
    (f5, v5), n5 = launches(lambda: m.predict_f(Xs))
    assert n5 == n_full
    fresh = gp.models.GPR((X, Y), kernel=k, noise_variance=0.05)
    f6, v6 = fresh.predict_f(Xs)
    assert np.array_equal(np.asarray(f5), np.asarray(f6)) and np.array_equal(np.asarray(v5), np.asarray(v6))
    assert not np.array_equal(np.asarray(f5), np.asarray(f1))
    # mean function: the vectors are recomputed from the re-centred targets even when W is reused
    mf = gp.mean_functions.Constant(0.3)
    mm = gp.models.GPR((X, Y), kernel=k, noise_variance=0.05, mean_function=mf)
    a1, _ = mm.predict_f(Xs)
    mf.c.assign(np.array([-0.4]))
    a2, _ = mm.predict_f(Xs)
    ref = gp.models.GPR((X, Y), kernel=k, noise_variance=0.05, mean_function=gp.mean_functions.Constant(-0.4)).predict_f(Xs)[0]
    np.testing.assert_allclose(np.asarray(a2), np.asarray(ref), rtol=1e-13, atol=1e-13)
    assert not np.allclose(np.asarray(a1), np.asarray(a2))


def test_evaluations_are_bitwise_reproducible(gp):
    """Fixed-order reductions, no atomics on values: repeated evaluations (side streams, PDL, the batched
    kernel's warp-specialised Cholesky) return identical bits."""
    X, Y = make_multi_input(3, 600, 8)
    k = gp.kernels.SquaredExponential() + gp.kernels.Matern52() + gp.kernels.Linear()
    m = gp.models.GPR((X, Y), kernel=k, noise_variance=1e-2)
    ref = m.lml_and_constrained_grads()
    for _ in range(40):
        out = m.lml_and_constrained_grads()
        assert out[0] == ref[0] and np.array_equal(out[1], ref[1]) and out[2] == ref[2]
    Xb = np.stack([X[i:i + 128] for i in range(40)])
    Yb = np.stack([Y[i:i + 128, 0] for i in range(40)])
    b = gp.BatchedGPR(Xb, Yb, k, noise_variance=0.1)
    r0 = b.lml_and_grads()
    for _ in range(20):
        assert all(np.array_equal(a, c) for a, c in zip(b.lml_and_grads(), r0))


@pytest.mark.parametrize("N", [700, 1024, 2500])
def test_factor_only_flows_match_the_full_inverse_flows(gp, N):
    """log_marginal_likelihood() and a cold predict_f take the factor-only path (N^3/3 flop, block forward
    substitution); the objective + gradient evaluation keeps W = L^-1.  Same numbers either way, and both
    against the oracle; a predict after either kind of evaluation reuses what the engine holds."""
    X, Y = make_multi_input(71, N, 4)
    Xs, _ = make_multi_input(72, 333, 4)
    k = gp.kernels.SquaredExponential(lengthscales=1.2) + gp.kernels.Matern52(variance=0.6, lengthscales=2.0)
    noise = 1e-2
    ko = to_oracle(k)
    l0 = O.gpr_lml(ko, X, Y, noise)
    m0, v0 = O.gpr_predict_f(ko, X, Y, noise, Xs)
    m = gp.models.GPR((X, Y), kernel=k, noise_variance=noise)
    eng = m._get_engine()
    mean_cold, var_cold = m.predict_f(Xs)                      # cold: factor only
    n0 = eng.launch_count()
    mean_again, var_again = m.predict_f(Xs)                    # reuse of the stored factor
    n_reuse = eng.launch_count() - n0
    lml_only = float(m.log_marginal_likelihood())              # value only: factor only
    lml, g, gn = m.lml_and_constrained_grads()                 # factor + inverse
    mean_w, var_w = m.predict_f(Xs)                            # reuse of W
    assert abs(lml_only - l0) <= 1e-9 * abs(l0) and abs(lml - l0) <= 1e-9 * abs(l0)
    assert abs(lml_only - lml) <= 1e-12 * abs(lml)
    for mean, var in ((mean_cold, var_cold), (mean_again, var_again), (mean_w, var_w)):
        assert np.max(np.abs(mean.numpy() - m0)) <= 1e-9 * np.max(np.abs(m0))
        assert np.max(np.abs(var.numpy() - v0) / np.abs(v0)) <= 1e-8
    assert np.array_equal(mean_cold.numpy(), mean_again.numpy()) and np.array_equal(var_cold.numpy(), var_again.numpy())
    m2 = gp.models.GPR((X, Y), kernel=k, noise_variance=noise * 1.5)
    n0 = eng.launch_count()
    m2.predict_f(Xs)
    assert n_reuse < eng.launch_count() - n0                   # the reuse skipped assembly and factorisation


@pytest.mark.parametrize("N", [3072, 3500, 5000])
def test_pipelined_factorisation_matches_the_recursion_and_the_oracle(gp, N):
    """Option 4 = 1, N >= 3072: objective + gradient run the right-looking pipeline over the two SM partitions
    (csrc/cholesky.cu factor_inv_pipelined); the default is the single-partition recursion.  Both against
    the oracle at the north-star bars, and a prediction from the factorisation the pipeline leaves behind."""
    X, Y = make_multi_input(81, N, 8)
    Xs, _ = make_multi_input(82, 200, 8)
    k = gp.kernels.SquaredExponential(lengthscales=1.3) + gp.kernels.Matern52(variance=0.7, lengthscales=2.0) + gp.kernels.Linear(variance=0.2)
    noise = 1e-2
    ko = to_oracle(k)
    l0, g0, n0 = O.gpr_lml_and_grad(ko, X, Y, noise)
    m0, v0 = O.gpr_predict_f(ko, X, Y, noise, Xs)
    m = gp.models.GPR((X, Y), kernel=k, noise_variance=noise)
    eng = m._get_engine()
    out = {}
    for mode in (1, 0):
        eng.set_option(eng.OPTION_PIPELINE, mode)
        try:
            lml, g, gn = m.lml_and_constrained_grads()
            mean, var = m.predict_f(Xs)
        finally:
            eng.set_option(eng.OPTION_PIPELINE, 0)
        gscale = max(1.0, np.max(np.abs(g0)), abs(n0))
        assert abs(lml - l0) <= 1e-9 * abs(l0), (mode, lml, l0)
        assert max(np.max(np.abs(g - g0)), abs(gn - n0)) <= 1e-7 * gscale, mode
        assert np.max(np.abs(mean.numpy() - m0)) <= 1e-9 * np.max(np.abs(m0)), mode
        assert np.max(np.abs(var.numpy() - v0) / np.abs(v0)) <= 1e-8, mode
        out[mode] = (lml, g)
    assert abs(out[0][0] - out[1][0]) <= 1e-12 * abs(l0)
    # with the stream forks off the same task list runs on one stream: same numbers, bit for bit
    eng.set_option(eng.OPTION_FORK_STREAMS, 0)
    eng.set_option(eng.OPTION_PIPELINE, 1)
    try:
        lml_s, g_s, _ = m.lml_and_constrained_grads()
    finally:
        eng.set_option(eng.OPTION_FORK_STREAMS, 1)
        eng.set_option(eng.OPTION_PIPELINE, 0)
    assert lml_s == out[1][0] and np.array_equal(g_s, out[1][1])
