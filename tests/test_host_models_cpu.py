"""CPU: the GPflow-shaped host layer end to end against a fake engine (tests/fake_engine.py) that
answers C-ABI-level requests with the CPU oracle: kernel lowering round trip, GPR objective /
gradient chain rule / predict, the reference's ModelTrainer call pattern with SciPy, deepcopy."""
import copy

import numpy as np
import pytest
import scipy.optimize
import torch

import portfoliooptgp_b200 as gpflow
from oracle import gpflow_oracle as O
from portfoliooptgp_b200 import ops
from portfoliooptgp_b200.kernels import compile_kernel
from tests.fake_engine import FakeEngine, spec_to_oracle
from tests.helpers import kernel_zoo, make_multi_input, to_oracle


@pytest.fixture(autouse=True)
def _direct_form():
    O.set_distance_form("direct")   # Gram-form rounding depends on the BLAS path taken for sliced vs unsliced X
    yield
    O.set_distance_form("gram")


@pytest.fixture
def fake(monkeypatch):
    eng = FakeEngine()
    monkeypatch.setattr(ops, "cuda_device_index", lambda device=None: 0)
    monkeypatch.setattr(ops, "shared_engine", lambda device=None: eng)
    monkeypatch.setattr(ops, "sync_stream", lambda e: None)

    def to_device(a, device=None, ndim=None):
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64)))
        t = t.detach().to(torch.float64)
        if ndim == 2 and t.ndim == 1:
            t = t[:, None]
        return t.contiguous()

    monkeypatch.setattr(ops, "to_device", to_device)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self, raising=False)
    import portfoliooptgp_b200.models as M
    monkeypatch.setattr(M, "_out", lambda t: t)
    orig_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: orig_empty(*a, **{kk: vv for kk, vv in k.items() if kk != "device"}))
    return eng


@pytest.mark.parametrize("D", [1, 3, 8])
def test_lowered_spec_round_trips_to_the_same_kernel(D):
    """compile_kernel -> gpb_kernel_spec -> oracle tree evaluates to the same K as the direct translation."""
    X, _ = make_multi_input(5, 40, D)
    for name, k in kernel_zoo(D).items():
        ck = compile_kernel(k, D)
        rebuilt, _ = spec_to_oracle(ck.spec, ck.theta())
        assert np.allclose(O.K(rebuilt, X), O.K(to_oracle(k), X), rtol=1e-14, atol=1e-14), name


def test_gpr_objective_and_chain_rule(fake):
    X, Y = make_multi_input(7, 60, 3)
    for name, k in kernel_zoo(3).items():
        m = gpflow.models.GPR((X, Y), kernel=k, noise_variance=0.05)
        ko = to_oracle(k)
        l0, g0, n0 = O.gpr_lml_and_grad(ko, X, Y, 0.05)
        assert float(m.log_marginal_likelihood()) == pytest.approx(l0, rel=1e-12)
        assert float(m.training_loss()) == pytest.approx(-l0, rel=1e-12)
        variables = m.trainable_variables
        loss, grads = m.training_loss_closure().value_and_grads(variables)
        # finite differences of the loss in UNCONSTRAINED space check the softplus chain rule + scattering
        x0 = gpflow.optimizers.Scipy.initial_parameters(variables)
        g = gpflow.optimizers.Scipy.pack_tensors(grads)

        def f(x):
            gpflow.optimizers.Scipy.assign_tensors(variables, x)
            return float(m.training_loss())

        h = 1e-6
        for i in range(len(x0)):
            e = np.zeros_like(x0); e[i] = h
            fd = (f(x0 + e) - f(x0 - e)) / (2 * h)
            assert g[i] == pytest.approx(fd, rel=2e-5, abs=2e-6), (name, i)
        gpflow.optimizers.Scipy.assign_tensors(variables, x0)


def test_reference_model_trainer_pattern_on_cpu(fake):
    """GPR/model_trainer.py:10-26 with the import swapped, numerics served by the oracle."""
    X, Y = make_multi_input(9, 50, 1)
    kernels = [gpflow.kernels.SquaredExponential(), gpflow.kernels.Exponential() + gpflow.kernels.Linear()]
    best = None
    for kernel in kernels:
        model = gpflow.models.GPR(data=(X, Y), kernel=kernel)
        model.likelihood.variance.assign(1e-2)
        gpflow.set_trainable(model.likelihood.variance, False)
        before = float(model.training_loss())
        res = gpflow.optimizers.Scipy().minimize(model.training_loss, model.trainable_variables, options=dict(maxiter=25))
        assert res.fun <= before and float(model.training_loss()) == pytest.approx(res.fun, rel=1e-10)
        mean, var = model.predict_f(X)
        y_mean, y_var = model.predict_y(X)
        assert np.allclose(y_var.numpy() - var.numpy(), 1e-2)
        mse = float(np.mean((Y - mean.numpy()) ** 2))
        best = mse if best is None else min(best, mse)
    assert best < 1.0
    # frozen likelihood was not touched; trained kernel parameters moved
    assert float(model.likelihood.variance.numpy()) == pytest.approx(1e-2, rel=1e-12)
    assert float(kernels[0].lengthscales.numpy()) != 1.0


def test_train_likelihood_restarts_pattern_and_deepcopy(fake):
    """Multi-Input_GPR/models/model_trainer.py:26-54: restarts over the starting noise, keep min loss."""
    X, Y = make_multi_input(11, 40, 3)
    composite = gpflow.kernels.Exponential(active_dims=slice(0, 2)) * gpflow.kernels.Exponential(active_dims=slice(2, 3))
    best_loss, best_model = float("inf"), None
    for start_var in [1e-5, 1e-3, 1e-1, 1.0]:
        model = gpflow.models.GPR((X, Y), kernel=copy.deepcopy(composite), noise_variance=start_var)
        gpflow.set_trainable(model.likelihood, True)
        logs = gpflow.optimizers.Scipy().minimize(model.training_loss, model.trainable_variables, options=dict(maxiter=15))
        if logs.fun < best_loss:
            best_loss, best_model = logs.fun, model
    assert np.isfinite(best_loss) and best_model is not None
    assert float(composite.kernels[0].variance.numpy()) == 1.0     # the template kernel was deep-copied, not trained
    m2 = copy.deepcopy(best_model)
    assert float(m2.training_loss()) == pytest.approx(float(best_model.training_loss()), rel=1e-12)


def test_sgpr_host_layer_against_fake_engine(fake):
    """models.SGPR on the CPU: objective sign, variable order (Z, kernel, noise, mean function),
    softplus chain rule, the err = Y - m(X) sign of the mean-function gradient, predict_f / predict_y,
    and the reference's plot_model pattern (test_scripts/SVGP.py:393-399) with SciPy."""
    rng = np.random.default_rng(4)
    N, D, M = 80, 2, 9
    X, Y = make_multi_input(9, N, D)
    Z0 = X[rng.choice(N, M, replace=False)].copy()
    A0, b0 = np.array([[0.2], [-0.1]]), np.array([0.05])
    k = gpflow.kernels.SquaredExponential(variance=0.9, lengthscales=1.2)
    mf = gpflow.mean_functions.Linear(A=A0, b=b0)
    m = gpflow.models.SGPR((X, Y), kernel=k, inducing_variable=Z0, noise_variance=0.1, mean_function=mf)
    ko = to_oracle(k)
    mean = X @ A0 + b0
    want = O.sgpr_elbo(ko, Z0, 0.1, X, Y, mean=mean)
    assert float(m.elbo()) == pytest.approx(want, rel=1e-12)
    assert float(m.training_loss()) == pytest.approx(-want, rel=1e-12)
    tv = m.trainable_variables
    assert tv[0] is m.inducing_variable.Z.unconstrained_variable
    loss, grads = m.training_loss_closure().value_and_grads(tv)
    # finite differences of the oracle objective in UNCONSTRAINED space, variable by variable
    def loss_at():
        kk = to_oracle(k)
        mm = X @ mf.A.numpy() + mf.b.numpy()
        return -O.sgpr_elbo(kk, m.inducing_variable.Z.numpy(), float(m.likelihood.variance.numpy()), X, Y, mean=mm)
    for v, g in zip(tv, grads):
        flat = v._value.reshape(-1)
        for idx in (0, flat.size - 1):
            old = flat[idx]
            h = 1e-6
            flat[idx] = old + h; fp = loss_at()
            flat[idx] = old - h; fm_ = loss_at()
            flat[idx] = old
            fd = (fp - fm_) / (2 * h)
            assert np.ravel(g)[idx] == pytest.approx(fd, rel=2e-5, abs=2e-6)
    Xs = rng.normal(size=(11, D))
    fm, fv = m.predict_f(Xs)
    om, ov = O.sgpr_predict_f(ko, Z0, 0.1, X, Y, Xs, mean=mean)
    np.testing.assert_allclose(np.asarray(fm), om + Xs @ A0 + b0, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(np.asarray(fv), ov, rtol=1e-12, atol=1e-14)
    ym, yv = m.predict_y(Xs)
    np.testing.assert_allclose(np.asarray(yv), ov + 0.1, rtol=1e-12)
    res = gpflow.optimizers.Scipy().minimize(m.training_loss, m.trainable_variables, options=dict(maxiter=15))
    assert res.fun < loss
