"""The straight-line (StaticShape) instantiations of the element kernels (csrc/shapes.cuh) against the
run-time interpreter on the same inputs: same arithmetic, so the results must agree to rounding of
a re-ordered instruction stream (1e-13 relative); and against the CPU oracle through the ordinary
parity tests, which run with the shapes enabled.  Also checks which expressions are matched."""
import numpy as np
import pytest

from tests.helpers import make_multi_input

pytestmark = pytest.mark.gpu


def _shapes(gp, D):
    K = gp.kernels
    last = [D - 1]
    feat = slice(0, max(D - 1, 1))
    out = {
        "se": (1, K.SquaredExponential(variance=1.3, lengthscales=0.7)),
        "m12": (2, K.Matern12(variance=0.8, lengthscales=1.4)),
        "rq": (3, K.RationalQuadratic(variance=1.1, lengthscales=0.9, alpha=0.7)),
        "exp": (4, K.Exponential(variance=0.9, lengthscales=1.2)),
        "se+m12": (5, K.SquaredExponential(lengthscales=0.9) + K.Matern12(variance=0.5, lengthscales=2.0)),
        "se*m12": (6, K.SquaredExponential(variance=1.2, lengthscales=0.8) * K.Matern12(variance=0.5, lengthscales=2.0)),
        "se+m52+lin": (8, K.SquaredExponential(lengthscales=1.1) + K.Matern52(variance=0.6, lengthscales=1.7) + K.Linear(variance=0.2)),
        "se+m52": (9, K.SquaredExponential(lengthscales=1.1) + K.Matern52(variance=0.6, lengthscales=1.7)),
        "se+per(se)": (10, K.SquaredExponential(lengthscales=1.3) + K.Periodic(K.SquaredExponential(variance=0.5, lengthscales=0.9, active_dims=last), period=1.9)),
        "exp+per(se)": (11, K.Exponential(lengthscales=1.3) + K.Periodic(K.SquaredExponential(variance=0.5, lengthscales=0.9, active_dims=last), period=1.9)),
        # three groups (GPR/main.py:111)
        "exp+per(se)+lin": (12, K.Exponential(lengthscales=1.3) + K.Periodic(K.SquaredExponential(active_dims=last), period=1.9) + K.Linear(variance=0.3)),
        # not in the table: stays on the interpreter
        "m32": (0, K.Matern32(lengthscales=1.1)),
    }
    if D >= 2:
        out["exp*exp"] = (7, K.Exponential(lengthscales=1.5, active_dims=feat) * K.Exponential(variance=0.7, lengthscales=0.8, active_dims=last))
    return out


def _engine(gp):
    from portfoliooptgp_b200 import ops
    return ops.shared_engine()


@pytest.fixture
def restore_option(gp):
    eng = _engine(gp)
    yield eng
    eng.set_option(eng.OPTION_STATIC_SHAPES, 1)


@pytest.mark.parametrize("D", [1, 3, 8, 12])
def test_matching_and_gpr_equivalence(gp, restore_option, D):
    eng = restore_option
    X, Y = make_multi_input(31 + D, 300, D)
    Xs = np.random.default_rng(D).normal(size=(50, D))
    for name, (sid, k) in _shapes(gp, D).items():
        res = {}
        for on in (1, 0):
            eng.set_option(eng.OPTION_STATIC_SHAPES, on)
            m = gp.models.GPR((X, Y), kernel=k, noise_variance=0.05)
            loss, grads = m.training_loss_closure().value_and_grads(m.trainable_variables)
            if on:
                assert eng.kernel_shape() == sid, name
            else:
                assert eng.kernel_shape() == 0, name
            fm, fv = m.predict_f(Xs)
            Kd = np.asarray(gp.ops.kernel_matrix(k, X[:70], X[70:200]).cpu())
            res[on] = (loss, np.concatenate([np.ravel(g) for g in grads]), np.asarray(fm), np.asarray(fv), Kd)
        a, b = res[1], res[0]
        assert a[0] == pytest.approx(b[0], rel=1e-13), name
        np.testing.assert_allclose(a[1], b[1], rtol=1e-10, atol=1e-10 * np.max(np.abs(b[1])), err_msg=name)
        np.testing.assert_allclose(a[2], b[2], rtol=1e-11, atol=1e-12, err_msg=name)
        np.testing.assert_allclose(a[3], b[3], rtol=1e-10, atol=1e-12, err_msg=name)
        np.testing.assert_allclose(a[4], b[4], rtol=1e-14, atol=1e-15, err_msg=name)


@pytest.mark.parametrize("D", [1, 2, 8])
def test_batched_equivalence(gp, restore_option, D):
    eng = restore_option
    X, Y = make_multi_input(5 + D, 100 + 9, D)
    B, N = 10, 100
    Xb = np.stack([X[i:i + N] for i in range(B)])
    Yb = np.stack([Y[i:i + N, 0] for i in range(B)])
    Xn = np.random.default_rng(1).normal(size=(B, 7, D))
    for name, (sid, k) in _shapes(gp, D).items():
        res = {}
        for on in (1, 0):
            eng.set_option(eng.OPTION_STATIC_SHAPES, on)
            m = gp.BatchedGPR(Xb, Yb, k, noise_variance=0.1)
            f = m.lml_and_grads()
            pm, pv = m.predict_f(Xn)
            res[on] = (f, pm.cpu().numpy(), pv.cpu().numpy())
        for x, y in zip(res[1][0][:3], res[0][0][:3]):
            np.testing.assert_allclose(x, y, rtol=1e-10, atol=1e-10 * max(1.0, np.max(np.abs(y))), err_msg=name)
        assert (res[1][0][3] == 0).all() and (res[0][0][3] == 0).all(), name
        np.testing.assert_allclose(res[1][1], res[0][1], rtol=1e-10, atol=1e-11, err_msg=name)
        np.testing.assert_allclose(res[1][2], res[0][2], rtol=1e-10, atol=1e-11, err_msg=name)


def test_svgp_and_sgpr_equivalence(gp, restore_option):
    eng = restore_option
    D, M, N = 3, 30, 400
    X, Y = make_multi_input(77, N, D)
    Z = X[::14][:M].copy()
    for name, (sid, k) in _shapes(gp, D).items():
        res = {}
        for on in (1, 0):
            eng.set_option(eng.OPTION_STATIC_SHAPES, on)
            sv = gp.models.SVGP(kernel=k, likelihood=gp.likelihoods.Gaussian(variance=0.05), inducing_variable=Z.copy(), num_data=N)
            l1, g1 = sv.training_loss_closure((X, Y)).value_and_grads(sv.trainable_variables)
            sg = gp.models.SGPR((X, Y), kernel=k, inducing_variable=Z.copy(), noise_variance=0.05)
            l2, g2 = sg.training_loss_closure().value_and_grads(sg.trainable_variables)
            res[on] = (l1, np.concatenate([np.ravel(g) for g in g1]), l2, np.concatenate([np.ravel(g) for g in g2]))
        a, b = res[1], res[0]
        assert a[0] == pytest.approx(b[0], rel=1e-12), name
        assert a[2] == pytest.approx(b[2], rel=1e-12), name
        np.testing.assert_allclose(a[1], b[1], rtol=1e-9, atol=1e-9 * np.max(np.abs(b[1])), err_msg=name)
        np.testing.assert_allclose(a[3], b[3], rtol=1e-9, atol=1e-9 * np.max(np.abs(b[3])), err_msg=name)


def test_every_reference_candidate_has_a_straight_line_shape(gp):
    """The 8 kernel candidates of GPR/main.py:105-114 at the reference's D = 1: none of them runs on the
    interpreter (gpb_kernel_shape != 0), VERDICT r01 missing #8."""
    from portfoliooptgp_b200.kernels import compile_kernel
    K = gp.kernels
    cands = [K.SquaredExponential(), K.Matern12(), K.RationalQuadratic(), K.Exponential(), K.SquaredExponential() + K.Matern12(),
             K.Exponential() + K.Periodic(K.SquaredExponential()) + K.Linear(), K.Exponential() + K.Periodic(K.SquaredExponential()),
             K.SquaredExponential() * K.Matern12()]
    eng = _engine(gp)
    ids = []
    for k in cands:
        eng.set_kernel(compile_kernel(k, 1).spec)
        ids.append(eng.kernel_shape())
    assert all(i != 0 for i in ids) and len(set(ids)) == 8, ids
