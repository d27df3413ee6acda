"""GPU parity: SGPR (Titsias' collapsed bound; reference call site test_scripts/SVGP.py:393-399) -- ELBO,
its gradient w.r.t. kernel hyper-parameters, noise variance, inducing points and mean-function
parameters, predict_f / predict_y and a Scipy fit, vs the CPU oracles (numpy restatement of
gpflow/models/sgpr.py for values, torch autograd for gradients).
Tolerances: 1e-9 relative on ELBO / mean / variance, 1e-7 on gradients, 1e-6 on optimiser end points."""
import numpy as np
import pytest

from oracle import gpflow_oracle as O
from oracle import gpflow_oracle_torch as T
from tests.helpers import make_multi_input, to_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _direct_form():
    O.set_distance_form("direct")
    yield
    O.set_distance_form("gram")


def _kernels(gp, D):
    K = gp.kernels
    last = [D - 1]
    return {
        "se": K.SquaredExponential(variance=1.2, lengthscales=0.8),
        "matern52": K.Matern52(variance=0.9, lengthscales=1.3),
        "rq": K.RationalQuadratic(variance=1.1, lengthscales=1.2, alpha=0.8),
        "se+matern12+lin": K.SquaredExponential(lengthscales=0.9) + K.Matern12(variance=0.5, lengthscales=2.0)
        + K.Linear(variance=0.3),
        "exp*per(se)": K.Exponential(lengthscales=1.5) * K.Periodic(K.SquaredExponential(active_dims=last), period=2.3),
    }


def _setup(gp, k, M, N, D, seed=0, noise=0.05, mean_function=None):
    rng = np.random.default_rng(seed)
    X, Y = make_multi_input(80 + seed, N, D)
    Z = X[rng.choice(N, M, replace=False)] + 0.01 * rng.standard_normal((M, D))
    m = gp.models.SGPR((X, Y), kernel=k, inducing_variable=Z, noise_variance=noise, mean_function=mean_function)
    return m, X, Y, Z


def _theta_grads(k, by_var):
    got = []
    for p in k.parameters:
        u = p.unconstrained_variable.numpy()
        got.append(np.atleast_1d(-by_var[id(p.unconstrained_variable)] / p.transform.forward_grad(u)))
    return np.concatenate(got)


@pytest.mark.parametrize("M,N,D", [(10, 90, 1), (33, 257, 3), (150, 700, 8), (260, 1500, 4)])
def test_elbo_and_gradients_match_oracle(gp, M, N, D):
    for name, k in _kernels(gp, D).items():
        m, X, Y, Z = _setup(gp, k, M, N, D)
        ko = to_oracle(k)
        e_np = O.sgpr_elbo(ko, Z, 0.05, X, Y)
        e0, g0 = T.sgpr_elbo_and_grad(ko, Z, 0.05, X, Y)
        assert float(m.elbo()) == pytest.approx(e_np, rel=1e-9), name
        assert float(m.maximum_log_likelihood_objective()) == pytest.approx(e_np, rel=1e-9), name
        assert float(m.training_loss()) == pytest.approx(-e_np, rel=1e-9), name
        variables = m.trainable_variables
        loss, grads = m.training_loss_closure().value_and_grads(variables)
        assert loss == pytest.approx(-e0, rel=1e-9), name
        by_var = {id(v): g for v, g in zip(variables, grads)}
        scale = max(1.0, float(np.max(np.abs(g0["theta"]))))
        assert np.max(np.abs(_theta_grads(k, by_var) - g0["theta"])) <= 1e-7 * scale, name
        gZ = -by_var[id(m.inducing_variable.Z.unconstrained_variable)]
        assert np.max(np.abs(gZ - g0["Z"])) <= 1e-7 * max(1.0, np.max(np.abs(g0["Z"]))), name
        pv = m.likelihood.variance
        gn = -by_var[id(pv.unconstrained_variable)] / pv.transform.forward_grad(pv.unconstrained_variable.numpy())
        assert abs(float(gn) - g0["noise"]) <= 1e-7 * max(1.0, abs(g0["noise"])), name


def test_mean_function_gradients(gp):
    M, N, D = 25, 300, 2
    k = gp.kernels.SquaredExponential(variance=0.8, lengthscales=1.1)
    A0, b0 = np.array([[0.3], [-0.2]]), np.array([0.1])
    mf = gp.mean_functions.Linear(A=A0, b=b0)
    m, X, Y, Z = _setup(gp, k, M, N, D, seed=3, mean_function=mf)
    ko = to_oracle(k)
    mean = X @ A0 + b0
    assert float(m.elbo()) == pytest.approx(O.sgpr_elbo(ko, Z, 0.05, X, Y, mean=mean), rel=1e-9)
    e0, g0 = T.sgpr_elbo_and_grad(ko, Z, 0.05, X, Y, mean=mean)
    variables = m.trainable_variables
    loss, grads = m.training_loss_closure().value_and_grads(variables)
    by_var = {id(v): g for v, g in zip(variables, grads)}
    # err = Y - (X A + b): d elbo/dA = -X^T err_bar, d elbo/db = -sum err_bar
    gA = -by_var[id(mf.A.unconstrained_variable)]
    gb = -by_var[id(mf.b.unconstrained_variable)]
    wantA = -(X.T @ g0["err"][:, None])
    assert np.max(np.abs(gA - wantA)) <= 1e-7 * max(1.0, np.max(np.abs(wantA)))
    assert abs(float(gb[0]) + g0["err"].sum()) <= 1e-7 * max(1.0, abs(g0["err"].sum()))
    # predict_f adds the mean function back
    Xs = np.random.default_rng(5).normal(size=(40, D))
    fm, fv = m.predict_f(Xs)
    om, ov = O.sgpr_predict_f(ko, Z, 0.05, X, Y, Xs, mean=mean)
    np.testing.assert_allclose(np.asarray(fm), om + Xs @ A0 + b0, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(np.asarray(fv), ov, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("M,N,D,Ns", [(10, 90, 1, 200), (40, 500, 3, 1000), (150, 700, 8, 130)])
def test_predict_matches_oracle(gp, M, N, D, Ns):
    for name, k in _kernels(gp, D).items():
        m, X, Y, Z = _setup(gp, k, M, N, D, seed=1)
        Xs = np.random.default_rng(9).normal(size=(Ns, D))
        ko = to_oracle(k)
        om, ov = O.sgpr_predict_f(ko, Z, 0.05, X, Y, Xs)
        fm, fv = m.predict_f(Xs)
        np.testing.assert_allclose(np.asarray(fm), om, rtol=1e-9, atol=1e-9 * np.max(np.abs(om)), err_msg=name)
        np.testing.assert_allclose(np.asarray(fv), ov, rtol=1e-9, atol=1e-9 * np.max(np.abs(ov)), err_msg=name)
        ym, yv = m.predict_y(Xs)
        np.testing.assert_allclose(np.asarray(yv), ov + 0.05, rtol=1e-9, atol=1e-9 * np.max(np.abs(ov)), err_msg=name)
        assert np.array_equal(np.asarray(ym), np.asarray(fm))


def test_bound_is_tight_when_Z_equals_X(gp):
    """Titsias' bound <= exact LML with equality at Z = X (up to the Kuu jitter)."""
    X, Y = make_multi_input(7, 120, 2)
    k = gp.kernels.SquaredExponential(variance=1.1, lengthscales=0.9)
    lml = float(gp.models.GPR((X, Y), kernel=k, noise_variance=0.1).log_marginal_likelihood())
    tight = float(gp.models.SGPR((X, Y), kernel=k, inducing_variable=X.copy(), noise_variance=0.1).elbo())
    loose = float(gp.models.SGPR((X, Y), kernel=k, inducing_variable=X[:10].copy(), noise_variance=0.1).elbo())
    assert abs(tight - lml) < 1e-3 * abs(lml) and tight <= lml + 1e-6
    assert loose < tight


def test_reference_call_pattern_scipy_fit(gp):
    """test_scripts/SVGP.py:391-399 + plot_model: SGPR(data, SquaredExponential(), inducing_variable=linspace)
    -> Scipy().minimize(training_loss, trainable_variables) -> predict_y.  The same L-BFGS-B run on the
    CPU oracle's objective must end at the same point."""
    import scipy.optimize
    rng = np.random.default_rng(2)
    N = 300
    X = np.sort(rng.uniform(0, 360, size=(N, 1)), axis=0) / 100.0
    Y = np.sin(2.0 * X) + 0.3 * np.cos(5.0 * X) + 0.1 * rng.normal(size=(N, 1))
    Z0 = np.linspace(0, 3.6, 10)[:, None]
    k = gp.kernels.SquaredExponential()
    m = gp.models.SGPR((X, Y), kernel=k, inducing_variable=Z0.copy())
    tv = m.trainable_variables                                   # Z, lengthscales, variance, noise (tf.Module order)
    assert len(tv) == 4 and tv[0] is m.inducing_variable.Z.unconstrained_variable
    assert tv[3] is m.likelihood.variance.unconstrained_variable
    opt = gp.optimizers.Scipy()
    res = opt.minimize(m.training_loss, m.trainable_variables)
    assert res.success or res.nit > 5

    # oracle-side fit: same unconstrained packing [Z (10), lengthscales, variance, noise]
    sp, spi = O.softplus, O.softplus_inverse
    lower = 1e-6

    def f(u):
        Z = u[:10].reshape(10, 1)
        ls, var, nv = sp(u[10]), sp(u[11]), lower + sp(u[12])
        ko = O.Leaf("se", variance=var, lengthscales=ls)
        e, g = T.sgpr_elbo_and_grad(ko, Z, nv, X, Y)
        sig = O.sigmoid
        # oracle theta order for a Leaf: query it rather than assume
        th_names = [n for n, _, _ in O.kernel_params(ko)]
        gl = g["theta"][[i for i, n in enumerate(th_names) if "lengthscales" in n][0]]
        gv = g["theta"][[i for i, n in enumerate(th_names) if "variance" in n][0]]
        grad = np.concatenate([g["Z"].reshape(-1), [gl * sig(u[10]), gv * sig(u[11]), g["noise"] * sig(u[12])]])
        return -e, -grad

    u0 = np.concatenate([Z0.reshape(-1), [spi(1.0), spi(1.0), spi(1.0 - lower)]])
    ref = scipy.optimize.minimize(f, u0, jac=True, method="L-BFGS-B")
    assert float(m.training_loss()) == pytest.approx(ref.fun, rel=1e-6)
    np.testing.assert_allclose(float(k.lengthscales.numpy()), sp(ref.x[10]), rtol=1e-4)
    np.testing.assert_allclose(float(m.likelihood.variance.numpy()), lower + sp(ref.x[12]), rtol=1e-4)
    Xplot = np.linspace(0.0, 3.6, 200)[:, None]
    ym, yv = m.predict_y(Xplot)
    ko = O.Leaf("se", variance=float(k.variance.numpy()), lengthscales=float(k.lengthscales.numpy()))
    om, ov = O.sgpr_predict_f(ko, m.inducing_variable.Z.numpy(), float(m.likelihood.variance.numpy()), X, Y, Xplot)
    np.testing.assert_allclose(np.asarray(ym), om, rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(np.asarray(yv), ov + float(m.likelihood.variance.numpy()), rtol=1e-8, atol=1e-10)


def test_cholesky_failure_is_reported(gp):
    X, Y = make_multi_input(4, 60, 1)
    k = gp.kernels.Linear(variance=1.0)     # rank-1 Kuu: jitter 1e-6 is what keeps it factorisable
    Z = np.repeat(X[:1], 12, axis=0)
    m = gp.models.SGPR((X, Y), kernel=k, inducing_variable=Z, noise_variance=0.1)
    try:
        v = float(m.elbo())
        assert np.isfinite(v)
    except gp.CholeskyError as e:
        assert "Cholesky decomposition was not successful" in str(e)
