"""The reference's ModelTrainer classes with their independent fits in flight at once
(portfoliooptgp_b200/trainers.py; GPR/model_trainer.py:6-26, Multi-Input_GPR/models/model_trainer.py:26-72):
every fit is the reference's own SciPy run on its own model, so a concurrent run must reproduce the
sequential loop bit for bit -- same selected kernel, same MSE / loss, same fitted parameters."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_call_pattern.npz"))


def _candidates(K):
    # GPR/main.py:105-114
    return [K.SquaredExponential(), K.Matern12(), K.RationalQuadratic(), K.Exponential(), K.SquaredExponential() + K.Matern12(),
            K.Exponential() + K.Periodic(K.SquaredExponential()) + K.Linear(), K.Exponential() + K.Periodic(K.SquaredExponential()),
            K.SquaredExponential() * K.Matern12()]


def _params(model):
    return np.concatenate([np.atleast_1d(p.numpy()).reshape(-1) for p in model.kernel.parameters])


def test_gpr_model_trainer_concurrent_equals_the_sequential_loop(gp):
    X, Y = G["aapl_d_X"], G["aapl_d_Y"]          # the reference's own AAPL daily series, N = 89
    # the reference's loop, verbatim apart from the import (GPR/model_trainer.py:10-26)
    best = (None, float("inf"), None)
    seq_params = []
    for kernel in _candidates(gp.kernels):
        model = gp.models.GPR(data=(X, Y), kernel=kernel)
        model.likelihood.variance.assign(1e-5)
        gp.set_trainable(model.likelihood.variance, False)
        gp.optimizers.Scipy().minimize(model.training_loss, model.trainable_variables, options=dict(maxiter=100))
        mean_test, _ = model.predict_f(X)
        mse = float(np.mean((Y - mean_test.numpy()) ** 2))
        seq_params.append(_params(model))
        if mse < best[1]:
            best = (kernel, mse, model)
    cands = _candidates(gp.kernels)
    bk, bmse, bmodel = gp.GPRModelTrainer(cands, max_workers=4).train_model(X, Y)
    assert type(bk) is type(best[0]) and bmse == best[1]
    for k, want in zip(cands, seq_params):
        got = np.concatenate([np.atleast_1d(p.numpy()).reshape(-1) for p in k.parameters])
        assert np.array_equal(got, want)
    # the selected model still predicts from the caller's thread (back on the shared handle)
    mean, var = bmodel.predict_f(X[:5])
    assert np.all(np.isfinite(mean.numpy())) and np.all(var.numpy() > 0)


def test_train_likelihood_restart_grid_concurrent_equals_sequential(gp):
    from copy import deepcopy
    from tests.helpers import make_multi_input
    X, Y = make_multi_input(23, 67, 7)           # Multi-Input_GPR sizes: N = 67, D = 7
    K = gp.kernels
    comp = K.Exponential(active_dims=slice(0, 6)) * K.Exponential(active_dims=slice(6, 7))     # main.py:126-135
    losses, params = [], []
    for start_var in (1e-5, 1e-3, 1e-1, 1.0):     # models/model_trainer.py:26-48, sequential
        m = gp.models.GPR((X, Y), kernel=deepcopy(comp), noise_variance=start_var)
        gp.set_trainable(m.likelihood, True)
        res = gp.optimizers.Scipy().minimize(m.training_loss, m.trainable_variables)
        losses.append(res.fun)
        params.append(np.concatenate([_params(m), [float(m.likelihood.variance.numpy())]]))
    best = gp.MultiInputModelTrainer.train_likelihood(X, Y, comp, summary=False)
    i = int(np.argmin(losses))
    assert np.array_equal(np.concatenate([_params(best), [float(best.likelihood.variance.numpy())]]), params[i])
    assert float(best.training_loss()) == pytest.approx(losses[i], rel=1e-12)


def test_models_sharing_a_kernel_are_fitted_in_order(gp):
    """Candidates that share a kernel instance warm-start each other in the reference's loop (SURVEY.md section 3
    aliasing note): they must not run at the same time."""
    from tests.helpers import make_multi_input
    X, Y = make_multi_input(29, 120, 1)
    shared = gp.kernels.SquaredExponential()
    def build():
        ms = [gp.models.GPR((X, Y * s), kernel=shared, noise_variance=1e-2) for s in (1.0, 1.5)]
        ms.append(gp.models.GPR((X, Y), kernel=gp.kernels.Matern12(), noise_variance=1e-2))
        return ms
    ms = build()
    out = gp.fit_concurrently(ms, options=dict(maxiter=20))
    after_conc = _params(ms[0])
    shared.lengthscales.assign(1.0); shared.variance.assign(1.0)
    ms2 = build()
    for m in ms2:
        gp.optimizers.Scipy().minimize(m.training_loss, m.trainable_variables, options=dict(maxiter=20))
    assert np.array_equal(after_conc, _params(ms2[0]))
    assert out[2][0].nit > 0
