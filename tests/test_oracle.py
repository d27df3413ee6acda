"""CPU: pin the oracle (parity is otherwise unpinned -- GPflow 2.9.1 is not installable here).
Known-answer tests from closed forms, the two independent oracles against each
other, extended-precision finite differences, and the SVGP <-> GPR identity."""
import math

import mpmath as mp
import numpy as np
import pytest

from oracle import gpflow_oracle as O
from oracle import gpflow_oracle_torch as T


def test_kernel_known_answers():
    x0 = np.zeros((1, 2)); x1 = np.array([[1.0, 1.0]])           # r2 = 2
    assert O.K(O.Leaf("se", 1.7), x0, x1)[0, 0] == pytest.approx(1.7 * math.exp(-1.0), rel=1e-15)
    x1 = np.array([[1.0, 0.0]])                                   # r = 1
    assert O.K(O.Leaf("matern12", 0.9), x0, x1)[0, 0] == pytest.approx(0.9 * math.exp(-1.0), rel=1e-15)
    x2 = np.array([[2.0, 0.0]])                                   # r = 2
    assert O.K(O.Leaf("exponential", 0.9), x0, x2)[0, 0] == pytest.approx(0.9 * math.exp(-1.0), rel=1e-15)
    s3, s5 = math.sqrt(3), math.sqrt(5)
    assert O.K(O.Leaf("matern32", 1.0), x0, x1)[0, 0] == pytest.approx((1 + s3) * math.exp(-s3), rel=1e-15)
    assert O.K(O.Leaf("matern52", 1.0), x0, x1)[0, 0] == pytest.approx((1 + s5 + 5 / 3) * math.exp(-s5), rel=1e-15)
    assert O.K(O.Leaf("rq", 2.0, alpha=3.0), x0, x1)[0, 0] == pytest.approx(2.0 * (1 + 0.5 / 3.0) ** -3.0, rel=1e-15)
    # lengthscale: r = |x - x'| / l
    assert O.K(O.Leaf("matern12", 1.0, lengthscales=2.0), x0, x2)[0, 0] == pytest.approx(math.exp(-1.0), rel=1e-15)
    # k(x, x) = variance
    for kind in O.STATIONARY_KINDS:
        assert O.K(O.Leaf(kind, 1.3), x1)[0, 0] == pytest.approx(1.3, rel=1e-15)
        assert O.K_diag(O.Leaf(kind, 1.3), x1)[0] == 1.3
    # Linear
    a = np.array([[1.0, 2.0]]); b = np.array([[3.0, -1.0]])
    assert O.K(O.Leaf("linear", 0.5), a, b)[0, 0] == pytest.approx(0.5 * 1.0)
    assert O.K_diag(O.Leaf("linear", 0.5), a)[0] == pytest.approx(2.5)


def test_periodic_known_answers():
    p, l, v = 1.7, 0.8, 1.1
    k = O.Periodic(O.Leaf("se", v, l), p)
    x0 = np.zeros((1, 1))
    assert O.K(k, x0, np.array([[p]]))[0, 0] == pytest.approx(v, rel=1e-14)            # one full period
    assert O.K(k, x0, np.array([[p / 2]]))[0, 0] == pytest.approx(v * math.exp(-0.5 / l ** 2), rel=1e-14)
    km = O.Periodic(O.Leaf("matern12", v, l), p)                                        # K_r base: sum |sin|/l
    assert O.K(km, x0, np.array([[p / 2]]))[0, 0] == pytest.approx(v * math.exp(-1.0 / l), rel=1e-14)


def test_sum_product_active_dims():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((7, 3))
    a = O.Leaf("se", 1.2, 0.7, active_dims=slice(0, 2)); b = O.Leaf("exponential", 0.6, 1.3, active_dims=[2])
    Ka, Kb = O.K(a, X), O.K(b, X)
    assert np.array_equal(O.K(O.Sum([a, b]), X), Ka + Kb)
    assert np.array_equal(O.K(O.Product([a, b]), X), Ka * Kb)
    assert np.allclose(Ka, O.K(O.Leaf("se", 1.2, 0.7), X[:, :2]), rtol=0, atol=0)
    assert np.allclose(O.K_diag(O.Product([a, b]), X), 1.2 * 0.6)


def test_softplus_roundtrip():
    th = np.array([1e-5, 1e-3, 0.5, 1.0, 30.0, 800.0])
    u = O.softplus_inverse(th)
    assert np.allclose(O.softplus(u), th, rtol=1e-14)
    assert O.softplus_inverse(1.0) == pytest.approx(math.log(math.e - 1.0), rel=1e-15)
    h = 1e-6
    assert np.allclose((O.softplus(u + h) - O.softplus(u - h)) / (2 * h), O.sigmoid(u), rtol=1e-6)


def test_lml_closed_form_n1():
    y, v, s2 = 0.7, 1.3, 0.2
    got = O.gpr_lml(O.Leaf("se", v), np.zeros((1, 1)), np.array([[y]]), s2)
    assert got == pytest.approx(-0.5 * y * y / (v + s2) - 0.5 * math.log(2 * math.pi * (v + s2)), rel=1e-15)


def _kernels(D):
    return [
        O.Sum([O.Leaf("se", 1.3, 0.7), O.Periodic(O.Leaf("se", 0.8, 1.2, active_dims=[D - 1]), 1.7), O.Leaf("linear", 0.5)]),
        O.Product([O.Leaf("exponential", 1.1, 0.9, active_dims=slice(0, D - 1)),
                   O.Leaf("exponential", 0.7, 1.4, active_dims=slice(D - 1, D))]),
        O.Sum([O.Leaf("matern52", 1.3, 0.7), O.Leaf("rq", 0.9, 1.1, alpha=0.6), O.Leaf("matern32", 0.5, 2.0)]),
        O.Sum([O.Leaf("se", 1.3, np.linspace(0.7, 1.4, D)), O.Periodic(O.Leaf("matern32", 0.8, 1.2, active_dims=[0]), 1.7)]),
    ]


def test_numpy_and_torch_oracles_agree():
    rng = np.random.default_rng(1)
    X = rng.standard_normal((80, 3)); Y = rng.standard_normal((80, 1))
    for k in _kernels(3):
        l1, g1, n1 = O.gpr_lml_and_grad(k, X, Y, 0.1)
        l2, g2, n2 = T.gpr_lml_and_grad(k, X, Y, 0.1)
        assert l1 == pytest.approx(l2, rel=1e-13)
        # analytic trace identity (direct derivative formulas) vs autodiff through the Gram-form graph
        assert np.max(np.abs(g1 - g2)) < 1e-7 * max(1.0, np.max(np.abs(g2)))
        assert n1 == pytest.approx(n2, rel=1e-11)


def test_gram_vs_direct_distance_gap():
    rng = np.random.default_rng(2)
    X = rng.standard_normal((60, 8)); Y = rng.standard_normal((60, 1))
    k = O.Sum([O.Leaf("se"), O.Leaf("matern52"), O.Leaf("linear")])
    a = O.gpr_lml(k, X, Y, 1e-2)
    O.set_distance_form("direct")
    try:
        b = O.gpr_lml(k, X, Y, 1e-2)
    finally:
        O.set_distance_form("gram")
    assert abs(a - b) < 1e-9 * abs(a)


def _mp_lml(kind, X, y, variance, ls, noise):
    """SE / matern12 single-leaf LML in 50-digit arithmetic (independent of numpy/LAPACK)."""
    mp.mp.dps = 50
    n = len(y)
    Km = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            r2 = sum((mp.mpf(float(X[i, d])) - mp.mpf(float(X[j, d]))) ** 2 for d in range(X.shape[1])) / mp.mpf(ls) ** 2
            kv = mp.exp(-r2 / 2) if kind == "se" else mp.exp(-mp.sqrt(r2))
            Km[i, j] = mp.mpf(variance) * kv + (mp.mpf(noise) if i == j else 0)
    L = mp.cholesky(Km)
    yv = mp.matrix([float(v) for v in y])
    a = mp.lu_solve(L, yv)
    return -sum(a[i] ** 2 for i in range(n)) / 2 - mp.mpf(n) / 2 * mp.log(2 * mp.pi) - sum(mp.log(L[i, i]) for i in range(n))


@pytest.mark.parametrize("kind", ["se", "matern12"])
def test_lml_and_grad_vs_extended_precision(kind):
    rng = np.random.default_rng(3)
    X = rng.standard_normal((12, 2)); y = rng.standard_normal(12)
    v, l, s2 = 1.3, 0.9, 0.05
    ref = _mp_lml(kind, X, y, v, l, s2)
    lml, g, gn = O.gpr_lml_and_grad(O.Leaf(kind, v, l), X, y[:, None], s2)
    assert abs(lml - float(ref)) < 1e-12 * abs(float(ref))
    h = mp.mpf("1e-20")
    d_l = (_mp_lml(kind, X, y, v, mp.mpf(l) + h, s2) - _mp_lml(kind, X, y, v, mp.mpf(l) - h, s2)) / (2 * h)
    d_v = (_mp_lml(kind, X, y, mp.mpf(v) + h, l, s2) - _mp_lml(kind, X, y, mp.mpf(v) - h, l, s2)) / (2 * h)
    d_n = (_mp_lml(kind, X, y, v, l, mp.mpf(s2) + h) - _mp_lml(kind, X, y, v, l, mp.mpf(s2) - h)) / (2 * h)
    assert g[0] == pytest.approx(float(d_l), rel=1e-10)   # order: lengthscales, variance
    assert g[1] == pytest.approx(float(d_v), rel=1e-10)
    assert gn == pytest.approx(float(d_n), rel=1e-10)


def test_predict_y_minus_predict_f_is_noise():
    rng = np.random.default_rng(4)
    X = rng.standard_normal((30, 2)); Y = rng.standard_normal((30, 1)); Xs = rng.standard_normal((9, 2))
    k = _kernels(2)[0]
    mf, vf = O.gpr_predict_f(k, X, Y, 0.03, Xs)
    my, vy = O.gpr_predict_y(k, X, Y, 0.03, Xs)
    assert np.array_equal(mf, my) and np.allclose(vy - vf, 0.03, rtol=0, atol=1e-16)
    m, v = O.gpr_predict_f(O.Leaf("se", 1.0, 0.5), X, Y, 1e-6, X)
    assert np.all(v < 1e-6) and np.all(v > -1e-9)  # posterior variance at a training input is below the noise


def test_gauss_kl_known_answers():
    M = 6
    assert O.gauss_kl(np.zeros((M, 1)), np.eye(M)[None]) == pytest.approx(0.0, abs=1e-15)
    rng = np.random.default_rng(5)
    mu = rng.standard_normal((M, 1)); Lq = np.tril(rng.standard_normal((M, M))) + 2 * np.eye(M)
    S = Lq @ Lq.T
    want = 0.5 * (np.trace(S) + float(mu.T @ mu) - M - np.linalg.slogdet(S)[1])
    assert O.gauss_kl(mu, Lq[None]) == pytest.approx(want, rel=1e-13)
    G = rng.standard_normal((M, M)); Kp = G @ G.T + np.eye(M)
    want = 0.5 * (np.trace(np.linalg.solve(Kp, S)) + float(mu.T @ np.linalg.solve(Kp, mu)) - M
                  + np.linalg.slogdet(Kp)[1] - np.linalg.slogdet(S)[1])
    assert O.gauss_kl(mu, Lq[None], Kp) == pytest.approx(want, rel=1e-13)


def test_svgp_equals_gpr_at_optimal_q(monkeypatch):
    """Z = X, whitened, q at its optimum => ELBO = LML of GPR.  The identity is
    exact up to O(N jitter / sigma^2) (Kuu carries the jitter, Kuf does not), so the jitter is
    shrunk for this check."""
    monkeypatch.setattr(O, "DEFAULT_JITTER", 1e-11)
    rng = np.random.default_rng(6)
    N = 25
    X = rng.standard_normal((N, 2)); Y = rng.standard_normal((N, 1))
    k = O.Sum([O.Leaf("se", 1.2, 0.8), O.Leaf("matern32", 0.4, 1.5)])
    s2 = 0.1
    Kuu = O.K(k, X) + O.DEFAULT_JITTER * np.eye(N)
    Lm = np.linalg.cholesky(Kuu)
    # optimal q(u) for whitened v = Lm^-1 u: S_v = (I + Lm^T Lm / s2)^-1, m_v = S_v Lm^T y / s2
    Sv = np.linalg.inv(np.eye(N) + Lm.T @ Lm / s2)
    mv = Sv @ Lm.T @ Y / s2
    elbo = O.svgp_elbo(k, X, mv, np.linalg.cholesky(Sv)[None], s2, X, Y)
    Kfull = Kuu + s2 * np.eye(N)
    Lf = np.linalg.cholesky(Kfull)
    lml = float(O.multivariate_normal(Y, Lf)[0])
    assert elbo == pytest.approx(lml, rel=1e-8)
    # torch oracle evaluates the same ELBO
    import torch
    e2 = T.svgp_elbo(k, torch.tensor(O.get_theta(k)), torch.tensor(X), torch.tensor(mv),
                     torch.tensor(np.linalg.cholesky(Sv)[None]), torch.tensor(s2, dtype=torch.float64),
                     torch.tensor(X), torch.tensor(Y))
    assert float(e2) == pytest.approx(elbo, rel=1e-12)


def test_svgp_minibatch_scaling_and_unwhitened():
    rng = np.random.default_rng(7)
    X = rng.standard_normal((40, 1)); Y = rng.standard_normal((40, 1)); Z = np.linspace(-2, 2, 7)[:, None]
    k = O.Leaf("se", 1.0, 0.7)
    qm = rng.standard_normal((7, 1)) * 0.1; qs = (np.eye(7) * 0.5 + np.tril(rng.standard_normal((7, 7))) * 0.05)[None]
    full = O.svgp_elbo(k, Z, qm, qs, 0.1, X, Y, num_data=40)
    kl = O.gauss_kl(qm, qs)
    half = O.svgp_elbo(k, Z, qm, qs, 0.1, X[:20], Y[:20], num_data=40) + O.svgp_elbo(k, Z, qm, qs, 0.1, X[20:], Y[20:], num_data=40)
    assert 0.5 * (half + 2 * kl) - kl == pytest.approx(full, rel=1e-12)
    # un-whitened parameterisation of the same q gives the same ELBO
    Lm = np.linalg.cholesky(O.K(k, Z) + O.DEFAULT_JITTER * np.eye(7))
    assert O.svgp_elbo(k, Z, Lm @ qm, (Lm @ qs[0])[None], 0.1, X, Y, whiten=False) == pytest.approx(
        O.svgp_elbo(k, Z, qm, qs, 0.1, X, Y), rel=1e-9)


def test_sgpr_oracle_identities(monkeypatch):
    """SGPR (gpflow/models/sgpr.py restatement): NumPy and torch versions agree; the collapsed bound is
    below the exact LML and tight at Z = X; predict_f at Z = X is the exact GP posterior; the torch
    gradient matches central differences."""
    from oracle import gpflow_oracle_torch as T
    monkeypatch.setattr(O, "DEFAULT_JITTER", 1e-11)
    rng = np.random.default_rng(3)
    X = rng.normal(size=(40, 2))
    Y = np.sin(X[:, :1]) + 0.1 * rng.normal(size=(40, 1))
    k = O.Sum([O.Leaf("se", variance=1.2, lengthscales=0.8), O.Leaf("matern32", variance=0.4, lengthscales=1.5)])
    Z = X[:9].copy()
    e_np = O.sgpr_elbo(k, Z, 0.1, X, Y)
    e_t, g = T.sgpr_elbo_and_grad(k, Z, 0.1, X, Y)
    assert e_np == pytest.approx(e_t, rel=1e-12)
    lml = O.gpr_lml(k, X, Y, 0.1)
    tight = O.sgpr_elbo(k, X, 0.1, X, Y)
    assert e_np < lml and tight == pytest.approx(lml, rel=1e-8)
    Xs = rng.normal(size=(7, 2))
    fm, fv = O.sgpr_predict_f(k, X, 0.1, X, Y, Xs)
    gm, gv = O.gpr_predict_f(k, X, Y, 0.1, Xs)
    np.testing.assert_allclose(fm, gm, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(fv, gv, rtol=1e-6, atol=1e-9)
    h = 1e-6
    fd_noise = (O.sgpr_elbo(k, Z, 0.1 + h, X, Y) - O.sgpr_elbo(k, Z, 0.1 - h, X, Y)) / (2 * h)
    assert g["noise"] == pytest.approx(fd_noise, rel=1e-6)
    Zp, Zm = Z.copy(), Z.copy()
    Zp[2, 1] += h
    Zm[2, 1] -= h
    fd_z = (O.sgpr_elbo(k, Zp, 0.1, X, Y) - O.sgpr_elbo(k, Zm, 0.1, X, Y)) / (2 * h)
    assert g["Z"][2, 1] == pytest.approx(fd_z, rel=1e-5, abs=1e-7)
