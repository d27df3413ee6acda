"""GPU parity across input dimensions: every padded-dimension instantiation of the kernels
(DP = 1, 2, 4, 8, 16 <-> D = 1, 2, 3, 5, 12, 16) for assembly, GPR LML/grad/predict, the batched path
and SVGP, including the Gram/DMMA assembly variant (D >= 3) and ARD lengthscales."""
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
    ks = {
        "se+m52+lin": K.SquaredExponential(lengthscales=1.3) + K.Matern52(variance=0.6, lengthscales=2.0) + K.Linear(variance=0.3),
        "exp": K.Exponential(variance=0.9, lengthscales=1.7),
        "ard_m32": K.Matern32(variance=1.1, lengthscales=np.linspace(0.9, 2.1, D)),
    }
    if D >= 2:
        ks["se[0:D-1]*m12[D-1]"] = (K.SquaredExponential(active_dims=slice(0, D - 1), lengthscales=1.5)
                                    * K.Matern12(active_dims=[D - 1], lengthscales=2.5))
    return ks


@pytest.mark.parametrize("D", [1, 2, 3, 5, 12, 16])
def test_all_paths_across_dimensions(gp, D):
    from portfoliooptgp_b200 import ops
    N, Ns = 150, 40
    X, Y = make_multi_input(70 + D, N, D)
    Xs, _ = make_multi_input(170 + D, Ns, D)
    noise = 2e-2
    for name, k in _kernels(gp, D).items():
        ko = to_oracle(k)
        # assembly: lower / symmetric / cross
        Kref = O.K(ko, X)
        for mode in (1, 2):
            got = ops.kernel_matrix(k, X, mode=mode).cpu().numpy()
            want = Kref
            if mode == 1:
                got, want = np.tril(got), np.tril(want)
            assert np.max(np.abs(got - want)) < 2e-13, (name, D, mode)
        assert np.max(np.abs(ops.kernel_matrix(k, X, Xs).cpu().numpy() - O.K(ko, X, Xs))) < 2e-13, (name, D)
        # exact GP
        m = gp.models.GPR((X, Y), kernel=k, noise_variance=noise)
        lml, g, gn = m.lml_and_constrained_grads()
        l0, g0, n0 = O.gpr_lml_and_grad(ko, X, Y, noise)
        assert abs(lml - l0) <= 1e-9 * abs(l0), (name, D)
        assert np.max(np.abs(g - g0)) <= 1e-7 * max(1.0, np.max(np.abs(g0))), (name, D)
        assert abs(gn - n0) <= 1e-7 * max(1.0, abs(n0)), (name, D)
        mean, var = m.predict_f(Xs)
        m0, v0 = O.gpr_predict_f(ko, X, Y, noise, Xs)
        assert np.max(np.abs(mean.numpy() - m0)) <= 1e-9 * max(1.0, np.max(np.abs(m0))), (name, D)
        assert np.max(np.abs(var.numpy() - v0)) <= 1e-9, (name, D)
        # batched (two windows of 100 rows)
        Xb = np.stack([X[:100], X[50:150]]); Yb = np.stack([Y[:100, 0], Y[50:150, 0]])
        bm = gp.BatchedGPR(Xb, Yb, k, noise_variance=noise)
        bl, bg, bn, info = bm.lml_and_grads()
        for b in range(2):
            lb, gb, nb_ = O.gpr_lml_and_grad(ko, Xb[b], Yb[b][:, None], noise)
            assert info[b] == 0 and abs(bl[b] - lb) <= 1e-9 * abs(lb), (name, D, b)
            assert np.max(np.abs(bg[b] - gb)) <= 1e-7 * max(1.0, np.max(np.abs(gb))), (name, D, b)
            assert abs(bn[b] - nb_) <= 1e-7 * max(1.0, abs(nb_)), (name, D, b)


@pytest.mark.parametrize("D", [2, 5, 12])
def test_svgp_across_dimensions(gp, D):
    rng = np.random.default_rng(D)
    N, M = 220, 30
    X, Y = make_multi_input(90 + D, N, D)
    Z = X[rng.choice(N, M, replace=False)] + 0.02
    qmu = 0.2 * rng.standard_normal((M, 1))
    qs = (0.7 * np.eye(M) + 0.05 * np.tril(rng.standard_normal((M, M))))[None]
    K = gp.kernels
    for name, k in {"se+lin": K.SquaredExponential(lengthscales=1.4) + K.Linear(variance=0.2),
                    "m52*exp": K.Matern52(active_dims=slice(0, D - 1), lengthscales=1.8) * K.Exponential(active_dims=[D - 1])}.items():
        m = gp.models.SVGP(kernel=k, likelihood=gp.likelihoods.Gaussian(variance=5e-2), inducing_variable=Z,
                           num_data=1000, q_mu=qmu, q_sqrt=qs)
        ko = to_oracle(k)
        e0, g0 = T.svgp_elbo_and_grad(ko, Z, qmu, qs, 5e-2, X, Y, num_data=1000)
        variables = m.trainable_variables
        loss, grads = m.training_loss_closure((X, Y)).value_and_grads(variables)
        assert loss == pytest.approx(-e0, rel=1e-9), (name, D)
        by_var = {id(v): g for v, g in zip(variables, grads)}
        gZ = -by_var[id(m.inducing_variable.Z.unconstrained_variable)]
        assert np.max(np.abs(gZ - g0["Z"])) <= 1e-7 * max(1.0, np.max(np.abs(g0["Z"]))), (name, D)
        got_theta = np.concatenate([np.atleast_1d(-by_var[id(p.unconstrained_variable)]
                                                  / p.transform.forward_grad(p.unconstrained_variable.numpy())) for p in k.parameters])
        assert np.max(np.abs(got_theta - g0["theta"])) <= 1e-7 * max(1.0, np.max(np.abs(g0["theta"]))), (name, D)
