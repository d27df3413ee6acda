"""GPU parity: SVGP ELBO, its gradient w.r.t. every trainable (kernel, noise, Z, q_mu, q_sqrt) and
predict_f vs the CPU oracles (numpy restatement for values, torch autograd for gradients).
Tolerances: 1e-9 relative on ELBO / mean / variance, 1e-7 on gradients (north_star)."""
import numpy as np
import pytest

from oracle import gpflow_oracle as O
from oracle import gpflow_oracle_torch as T
from tests.helpers import make_multi_input, to_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _direct_form():
    O.set_distance_form("direct")   # the numpy oracle evaluates distances as the CUDA kernels do
    yield
    O.set_distance_form("gram")


def _kernels(gp, D):
    K = gp.kernels
    last = [D - 1]
    ks = {
        "se": K.SquaredExponential(variance=1.2, lengthscales=0.8),
        "matern12": K.Matern12(variance=0.9, lengthscales=1.3),
        "rq": K.RationalQuadratic(variance=1.1, lengthscales=1.2, alpha=0.8),
        "se+matern12": K.SquaredExponential(lengthscales=0.9) + K.Matern12(variance=0.5, lengthscales=2.0),
        "exp+per(se)+lin": K.Exponential(variance=0.7, lengthscales=1.3)
        + K.Periodic(K.SquaredExponential(variance=0.6, lengthscales=1.1, active_dims=last), period=1.7) + K.Linear(variance=0.3),
        "se*matern12": K.SquaredExponential(variance=1.2, lengthscales=0.8) * K.Matern12(variance=0.5, lengthscales=2.0),
        "exp*per(se)": K.Exponential(lengthscales=1.5) * K.Periodic(K.SquaredExponential(active_dims=last), period=2.3),
    }
    return ks


def _setup(gp, k, M, N, D, seed=0, num_data=None, noise=1e-2):
    rng = np.random.default_rng(seed)
    X, Y = make_multi_input(60 + seed, N, D)
    Z = X[rng.choice(N, M, replace=False)] + 0.01 * rng.standard_normal((M, D))
    qmu = 0.3 * rng.standard_normal((M, 1))
    qs = (0.6 * np.eye(M) + 0.05 * np.tril(rng.standard_normal((M, M))))[None]
    m = gp.models.SVGP(kernel=k, likelihood=gp.likelihoods.Gaussian(variance=noise), inducing_variable=Z,
                       num_data=num_data, q_mu=qmu, q_sqrt=qs)
    return m, X, Y, Z, qmu, qs


@pytest.mark.parametrize("M,N,D,num_data", [(20, 90, 1, None), (33, 257, 3, 1000), (150, 700, 8, 700)])
def test_elbo_and_gradients_match_oracle(gp, M, N, D, num_data):
    for name, k in _kernels(gp, D).items():
        m, X, Y, Z, qmu, qs = _setup(gp, k, M, N, D, num_data=num_data)
        ko = to_oracle(k)
        e0, g0 = T.svgp_elbo_and_grad(ko, Z, qmu, qs, 1e-2, X, Y, num_data=num_data)
        assert float(m.elbo((X, Y))) == pytest.approx(O.svgp_elbo(ko, Z, qmu, qs, 1e-2, X, Y, num_data=num_data), rel=1e-9), name
        variables = m.trainable_variables
        loss, grads = m.training_loss_closure((X, Y)).value_and_grads(variables)
        assert loss == pytest.approx(-e0, rel=1e-9), name
        by_var = {id(v): g for v, g in zip(variables, grads)}
        scale = max(1.0, float(np.max(np.abs(g0["theta"]))), float(np.max(np.abs(g0["q_mu"]))))
        # kernel parameters: constrained gradient * softplus'(u), in oracle order
        got_theta = []
        for p in k.parameters:
            u = p.unconstrained_variable.numpy()
            got_theta.append(-by_var[id(p.unconstrained_variable)] / p.transform.forward_grad(u))
        assert np.max(np.abs(np.concatenate([np.atleast_1d(g) for g in got_theta]) - g0["theta"])) <= 1e-7 * scale, name
        gZ = -by_var[id(m.inducing_variable.Z.unconstrained_variable)]
        assert np.max(np.abs(gZ - g0["Z"])) <= 1e-7 * max(1.0, np.max(np.abs(g0["Z"]))), name
        gq = -by_var[id(m.q_mu.unconstrained_variable)]
        assert np.max(np.abs(gq - g0["q_mu"])) <= 1e-7 * scale, name
        gs = -by_var[id(m.q_sqrt.unconstrained_variable)]
        want = g0["q_sqrt"][:, np.tril_indices(M)[0], np.tril_indices(M)[1]]
        assert np.max(np.abs(gs - want)) <= 1e-7 * max(1.0, np.max(np.abs(want))), name
        pv = m.likelihood.variance
        gn = -by_var[id(pv.unconstrained_variable)] / pv.transform.forward_grad(pv.unconstrained_variable.numpy())
        assert abs(float(gn) - g0["noise"]) <= 1e-7 * max(1.0, abs(g0["noise"])), name


def test_predict_f_matches_oracle(gp):
    for name, k in _kernels(gp, 3).items():
        m, X, Y, Z, qmu, qs = _setup(gp, k, 40, 200, 3, seed=1)
        Xs, _ = make_multi_input(77, 301, 3)
        mean, var = m.predict_f(Xs)
        m0, v0 = O.svgp_predict_f(to_oracle(k), Z, qmu, qs, Xs)
        assert mean.shape == (301, 1) and var.shape == (301, 1)
        assert np.max(np.abs(mean.numpy() - m0)) <= 1e-9 * max(1.0, np.max(np.abs(m0))), name
        assert np.max(np.abs(var.numpy() - v0)) <= 1e-9 * max(1.0, np.max(np.abs(v0))), name


def test_prior_kl_and_defaults(gp):
    m = gp.models.SVGP(gp.kernels.SquaredExponential(), gp.likelihoods.Gaussian(1e-4), np.linspace(0, 3, 20)[:, None], num_data=50)
    assert float(m.prior_kl()) == pytest.approx(0.0, abs=1e-12)           # q = N(0, I): KL = 0
    names = [n for n, _ in m.named_parameters()]
    assert names == ["inducing_variable.Z", "kernel.lengthscales", "kernel.variance", "likelihood.variance", "q_mu", "q_sqrt"]
    assert m.inducing_variable.Z.shape == (20, 1) and m.q_sqrt.shape == (1, 20, 20)
    assert m.q_sqrt.unconstrained_variable.shape == (1, 210)


def test_reference_call_pattern_svgp(gp):
    """test_scripts/SVGP.py:515-540 with the import swapped."""
    gpflow = gp
    rng = np.random.default_rng(3)
    X = np.sort(rng.uniform(0, 6, size=(120, 1)), axis=0)
    Y = np.sin(X) + 0.1 * rng.standard_normal((120, 1))
    model = gpflow.models.SVGP(kernel=gpflow.kernels.SquaredExponential() + gpflow.kernels.Matern12(),
                               likelihood=gpflow.likelihoods.Gaussian(variance=1e-2),
                               inducing_variable=np.linspace(0, X.max(), 20)[:, None], num_data=len(X))
    gpflow.set_trainable(model.likelihood.variance, False)
    opt = gpflow.optimizers.Scipy()
    training_loss = model.training_loss_closure((X, Y))
    before = float(training_loss())
    res = opt.minimize(training_loss, model.trainable_variables, options=dict(maxiter=100))
    assert res.fun < before - 10.0
    mean, var = model.predict_f(X)
    assert float(np.mean((mean.numpy() - Y) ** 2)) < 0.05
    assert model.inducing_variable.Z.numpy().shape == (20, 1)


def test_data_parallel_trainer_single_rank(gp):
    """SVGPDataParallel on one rank: the ELBO it reports equals SVGP.elbo at the same parameters, and
    Adam steps increase it."""
    from portfoliooptgp_b200.svgp_dp import SVGPDataParallel
    X, Y = make_multi_input(91, 4096, 4)
    Z = X[:64].copy()
    k = gp.kernels.SquaredExponential(lengthscales=1.5)
    tr = SVGPDataParallel(k, 1e-1, Z, num_data=4096, X_shard=X, Y_shard=Y, minibatch_size=1024, lr=5e-2)
    e0 = tr.step(update=False)
    m = gp.models.SVGP(gp.kernels.SquaredExponential(lengthscales=1.5), gp.likelihoods.Gaussian(1e-1), Z, num_data=4096)
    assert e0 == pytest.approx(float(m.elbo((X[:1024], Y[:1024]))), rel=1e-12)
    tr.cursor = 0
    first = tr.step()
    vals = [tr.step() for _ in range(40)]
    assert np.mean(vals[-4:]) > first + 100.0
    assert np.allclose(np.triu(tr.q_sqrt.cpu().numpy(), 1), 0.0)


def test_data_parallel_trainer_trains_the_noise_when_asked(gp):
    """GPflow's likelihood variance is trainable by default; the trainer keeps it frozen unless
    train_noise=True (the reference freezes it, test_scripts/SVGP.py:524).  With train_noise the variance
    moves (through the softplus + 1e-6 transform) and the ELBO improves faster than with it frozen."""
    from portfoliooptgp_b200.svgp_dp import SVGPDataParallel
    X, Y = make_multi_input(92, 4096, 4)
    Z = X[:64].copy()
    runs = {}
    for train_noise in (False, True):
        k = gp.kernels.SquaredExponential(lengthscales=1.5)
        # the targets are z-scored (variance 1): a likelihood variance of 4 is far too large
        tr = SVGPDataParallel(k, 4.0, Z, num_data=4096, X_shard=X, Y_shard=Y, minibatch_size=1024, lr=5e-2,
                              train_noise=train_noise, seed=3)
        vals = [tr.step() for _ in range(80)]
        runs[train_noise] = (tr.noise, float(np.mean(vals[-4:])))
    assert runs[False][0] == 4.0
    assert 1e-6 < runs[True][0] < 2.5
    assert runs[True][1] > runs[False][1]
