"""Independent pin of the CPU oracle's exact-GP arithmetic: scikit-learn's GaussianProcessRegressor
(a separate implementation of Rasmussen & Williams Alg. 2.1, the algorithm behind gpflow.models.GPR)
must give the same log marginal likelihood, the same hyper-parameter gradient and the same predictive
moments for the kernel families the reference uses.  GPflow itself is not installable here (parity
unpinned, see oracle/gpflow_oracle.py); this ties the oracle to a third-party library that is.

Parametrisation map (GPflow -> sklearn): SquaredExponential(v, l) = C(v) * RBF(l); Matern12/32/52 =
C(v) * Matern(l, nu); Exponential(v, l) = exp(-r/2) = C(v) * Matern(2 l, 0.5); RationalQuadratic(v, l, a) =
C(v) * RationalQuadratic(l, a); Periodic(SE(v, l), p) = C(v) * ExpSineSquared(2 l, p); Linear(v) = C(v) *
DotProduct(sigma_0 -> 0); Gaussian noise = WhiteKernel.  sklearn differentiates w.r.t. log theta."""
import numpy as np
import pytest

from oracle import gpflow_oracle as O

sk = pytest.importorskip("sklearn.gaussian_process")
from sklearn.gaussian_process.kernels import (RBF, ConstantKernel as C, DotProduct, ExpSineSquared, Matern,  # noqa: E402
                                              RationalQuadratic, WhiteKernel)


@pytest.fixture(autouse=True)
def _direct_form():
    # sklearn evaluates distances by direct differences (scipy cdist); GPflow's Gram form differs from that
    # by its cancellation error (1e-9 relative on the LML for the kinked Matern12 -- the gap that
    # tests/test_oracle.py::test_gram_vs_direct_distance_gap bounds), so the comparison runs in direct form
    O.set_distance_form("direct")
    yield
    O.set_distance_form("gram")


def _data(n, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, d))
    y = np.sin(X[:, :1]) + 0.3 * X[:, -1:] + 0.1 * rng.normal(size=(n, 1))
    return X, y


# (name, oracle kernel, sklearn kernel, map from sklearn's free log-parameters to the oracle's theta order)
def _cases():
    out = []
    out.append(("se", O.Leaf("se", variance=1.3, lengthscales=0.7), C(1.3) * RBF(0.7)))
    out.append(("matern12", O.Leaf("matern12", variance=0.8, lengthscales=1.4), C(0.8) * Matern(1.4, nu=0.5)))
    out.append(("matern32", O.Leaf("matern32", variance=1.1, lengthscales=0.9), C(1.1) * Matern(0.9, nu=1.5)))
    out.append(("matern52", O.Leaf("matern52", variance=0.6, lengthscales=1.7), C(0.6) * Matern(1.7, nu=2.5)))
    out.append(("exponential", O.Leaf("exponential", variance=0.9, lengthscales=1.2), C(0.9) * Matern(2.4, nu=0.5)))
    out.append(("rq", O.Leaf("rq", variance=1.2, lengthscales=0.8, alpha=0.6), C(1.2) * RationalQuadratic(0.8, alpha=0.6)))
    out.append(("se+matern12", O.Sum([O.Leaf("se", variance=1.0, lengthscales=0.9), O.Leaf("matern12", variance=0.5, lengthscales=2.0)]),
                C(1.0) * RBF(0.9) + C(0.5) * Matern(2.0, nu=0.5)))
    out.append(("se*matern12", O.Product([O.Leaf("se", variance=1.2, lengthscales=0.8), O.Leaf("matern12", variance=0.5, lengthscales=2.0)]),
                (C(1.2) * RBF(0.8)) * (C(0.5) * Matern(2.0, nu=0.5))))
    out.append(("se+matern52+linear", O.Sum([O.Leaf("se", variance=1.0, lengthscales=1.1), O.Leaf("matern52", variance=0.6, lengthscales=1.7),
                                              O.Leaf("linear", variance=0.2)]),
                C(1.0) * RBF(1.1) + C(0.6) * Matern(1.7, nu=2.5) + C(0.2) * DotProduct(sigma_0=1e-9, sigma_0_bounds="fixed")))
    return out


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c[0])
def test_lml_gradient_and_prediction_match_sklearn(case):
    name, ko, ks = case
    X, y = _data(70, 3, seed=4)
    Xs = np.random.default_rng(9).normal(size=(25, 3))
    noise = 0.07
    gpr = sk.GaussianProcessRegressor(kernel=ks + WhiteKernel(noise), alpha=0.0, optimizer=None, normalize_y=False)
    gpr.fit(X, y)
    lml_sk, grad_sk = gpr.log_marginal_likelihood(gpr.kernel_.theta, eval_gradient=True)
    lml, g, gn = O.gpr_lml_and_grad(ko, X, y, noise)
    assert lml == pytest.approx(lml_sk, rel=1e-10)
    # gradient: compare in log space.  sklearn orders its free parameters per product factor
    # (constant_value, then the stationary kernel's own in alphabetical order: alpha before length_scale),
    # the oracle per leaf (see O.kernel_params); match them by (leaf, attribute)
    theta = O.get_theta(ko)
    names = [(id(owner), attr) for _, owner, attr in O.kernel_params(ko)]
    glog = {k: float(g[i] * theta[i]) for i, k in enumerate(names)}
    want = []
    leaves = [ko] if isinstance(ko, O.Leaf) else list(ko.kernels)
    for leaf in leaves:
        want.append(glog[(id(leaf), "variance")])
        if leaf.kind == "rq":
            want.append(glog[(id(leaf), "alpha")])
        if leaf.kind != "linear":
            want.append(glog[(id(leaf), "lengthscales")])
    want.append(gn * noise)
    assert len(want) == len(grad_sk)
    np.testing.assert_allclose(np.array(want), grad_sk, rtol=2e-7, atol=2e-8 * max(1.0, np.max(np.abs(grad_sk))))
    mean_sk, std_sk = gpr.predict(Xs, return_std=True)
    fm, fv = O.gpr_predict_f(ko, X, y, noise, Xs)
    np.testing.assert_allclose(fm[:, 0], np.ravel(mean_sk), rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(fv[:, 0] + noise, std_sk ** 2, rtol=1e-7, atol=1e-10)


def test_periodic_matches_exp_sine_squared():
    X, y = _data(60, 1, seed=2)
    Xs = np.linspace(-2, 2, 17)[:, None]
    ko = O.Sum([O.Leaf("exponential", variance=0.7, lengthscales=1.3),
                O.Periodic(O.Leaf("se", variance=0.6, lengthscales=1.1), period=1.7)])
    ks = C(0.7) * Matern(2.6, nu=0.5) + C(0.6) * ExpSineSquared(length_scale=2.2, periodicity=1.7)
    noise = 0.05
    gpr = sk.GaussianProcessRegressor(kernel=ks + WhiteKernel(noise), alpha=0.0, optimizer=None).fit(X, y)
    lml_sk, grad_sk = gpr.log_marginal_likelihood(gpr.kernel_.theta, eval_gradient=True)
    lml, g, gn = O.gpr_lml_and_grad(ko, X, y, noise)
    assert lml == pytest.approx(lml_sk, rel=1e-10)
    theta = O.get_theta(ko)
    by = {(type(owner).__name__, attr, i): g[i] * theta[i] for i, (_, owner, attr) in enumerate(O.kernel_params(ko))}
    exp_leaf, per = ko.kernels
    names = [(id(owner), attr) for _, owner, attr in O.kernel_params(ko)]
    glog = {k: float(g[i] * theta[i]) for i, k in enumerate(names)}
    # sklearn order: C, Matern.length_scale, C, ExpSineSquared.length_scale, ExpSineSquared.periodicity, noise
    want = [glog[(id(exp_leaf), "variance")], glog[(id(exp_leaf), "lengthscales")], glog[(id(per.base), "variance")],
            glog[(id(per.base), "lengthscales")], glog[(id(per), "period")], gn * noise]
    np.testing.assert_allclose(np.array(want), grad_sk, rtol=2e-7, atol=2e-8 * max(1.0, np.max(np.abs(grad_sk))))
    mean_sk = gpr.predict(Xs)
    fm, _ = O.gpr_predict_f(ko, X, y, noise, Xs)
    np.testing.assert_allclose(fm[:, 0], np.ravel(mean_sk), rtol=1e-8, atol=1e-10)
