"""CPU: the upsampling mirror reproduces pandas' reindex + interpolate('linear') (GPR/predictor.py:35-51)."""
import numpy as np
import pandas as pd

from portfoliooptgp_b200.postprocess import predict_combined, upsample_predictions


def _pandas(X_daily, X, pred):
    s = pd.Series(pred.reshape(-1), index=X.reshape(-1))
    return s.reindex(X_daily.reshape(-1)).interpolate(method="linear").values.reshape(-1, 1)


def test_upsample_matches_pandas():
    rng = np.random.default_rng(0)
    xd = np.array([0, 1, 4, 5, 6, 7, 8, 11, 12, 13, 14, 15, 18, 19, 20], dtype=np.float64)[:, None]   # trading days
    for xw in (xd[[1, 4, 8, 13]], xd[[0, 5, 10, 14]], xd[[3, 9]]):
        pw = rng.standard_normal((len(xw), 1))
        got = upsample_predictions(xd, xw, pw, period="w")
        want = _pandas(xd, xw, pw)
        assert np.allclose(got, want, rtol=0, atol=1e-15, equal_nan=True)
    assert upsample_predictions(xd, xd, xd, period="d") is xd


def test_predict_combined_blend():
    xd = np.arange(10.0)[:, None]; xw = xd[[0, 5, 9]]; xm = xd[[0, 9]]
    d = tuple(np.full((10, 1), v) for v in (1.0, 2.0, 3.0, 4.0))
    w = tuple(np.full((3, 1), v) for v in (10.0, 20.0, 30.0, 40.0))
    m = tuple(np.full((2, 1), v) for v in (100.0, 200.0, 300.0, 400.0))
    out = predict_combined(0.5, 0.3, d, w, m, xd, xw, xm)
    assert np.allclose(out[0], 0.5 * 1 + 0.3 * 10 + 0.2 * 100) and np.allclose(out[3], 0.5 * 4 + 0.3 * 40 + 0.2 * 400)


def test_blend_weight_optimizer_matches_reference_formulation():
    """GPR/optimizer.py:13-28 restated with sklearn + SLSQP here; same optimum and same loss."""
    from scipy.optimize import minimize
    from sklearn.metrics import mean_squared_error

    from portfoliooptgp_b200.postprocess import Optimizer

    rng = np.random.default_rng(11)
    n = 200
    d, w, m = (rng.normal(size=(n, 1)) for _ in range(3))
    Y = 0.55 * d + 0.3 * w + 0.15 * m + 0.01 * rng.normal(size=(n, 1))
    opt = Optimizer(lambda_=0.01)
    got = opt.optimize_weights(Y, d, w, m)

    def ref_loss(x):
        a, b = x
        return mean_squared_error(Y, a * d + b * w + (1 - a - b) * m) + 0.01 * (np.abs(a) + np.abs(b))

    ref = minimize(ref_loss, [0.33, 0.33], bounds=[(0, 1), (0, 1)],
                   constraints={"type": "ineq", "fun": lambda x: 1 - sum(x)}, method="SLSQP").x
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-9)
    assert abs(opt.loss_fn(got, Y, d, w, m) - ref_loss(ref)) < 1e-12
    assert 0 <= got[0] <= 1 and 0 <= got[1] <= 1 and got.sum() <= 1 + 1e-9
    assert abs(got[0] - 0.55) < 0.05 and abs(got[1] - 0.3) < 0.05


# ---- device path (csrc/prep.cu gpb_post_upsample / gpb_post_blend) ------------------------------------------------
import pytest  # noqa: E402


def _grids(n_daily, seed):
    """A daily grid with weekend-like gaps, weekly / monthly sub-grids (a few weekly values off the grid)."""
    rng = np.random.default_rng(seed)
    days = np.cumsum(rng.choice([1.0, 1.0, 1.0, 1.0, 3.0], size=n_daily))
    xw = days[2::5].copy()
    xw[::17] += 0.5                      # not daily grid values: reindex drops them
    xm = days[7::21].copy()
    return days[:, None], np.unique(xw)[:, None], xm[:, None]


@pytest.mark.gpu
@pytest.mark.parametrize("n_daily", [15, 1000, 1_200_000])
def test_device_upsample_and_blend_match_pandas_bit_for_bit(n_daily):
    import torch
    from portfoliooptgp_b200.postprocess import upsample_predictions_device
    xd, xw, xm = _grids(n_daily, n_daily)
    rng = np.random.default_rng(1)
    pw = [rng.standard_normal((len(xw), 1)) for _ in range(4)]
    pm = [rng.standard_normal((len(xm), 1)) for _ in range(4)]
    pd_ = [rng.standard_normal((len(xd), 1)) for _ in range(4)]
    cu = lambda a: torch.as_tensor(a, device="cuda")
    got_w = upsample_predictions_device(cu(xd), cu(xw), tuple(cu(p) for p in pw), period="w")
    for q in range(4):
        want = _pandas(xd, xw, pw[q])
        assert np.array_equal(got_w[q].cpu().numpy(), want, equal_nan=True), q
    single = upsample_predictions(cu(xd), cu(xm), cu(pm[0]), period="m")       # public entry point, CUDA input
    assert np.array_equal(single.cpu().numpy(), _pandas(xd, xm, pm[0]), equal_nan=True)
    out = predict_combined(0.45, 0.35, tuple(cu(p) for p in pd_), tuple(cu(p) for p in pw), tuple(cu(p) for p in pm),
                           cu(xd), cu(xw), cu(xm))
    for q in range(4):
        want = 0.45 * pd_[q] + 0.35 * _pandas(xd, xw, pw[q]) + (1 - 0.45 - 0.35) * _pandas(xd, xm, pm[q])
        assert np.array_equal(out[q].cpu().numpy(), want, equal_nan=True), q
    host = predict_combined(0.45, 0.35, tuple(pd_), tuple(pw), tuple(pm), xd, xw, xm)
    for q in range(4):
        assert np.array_equal(out[q].cpu().numpy(), host[q], equal_nan=True)


@pytest.mark.gpu
def test_predictor_combines_three_models_on_the_device(gp):
    """GPR/predictor.py:10-33 end to end with device outputs: three fitted-size models, predict_single on each,
    upsample + blend without leaving the GPU; same numbers as the host path."""
    from portfoliooptgp_b200 import models
    from portfoliooptgp_b200.postprocess import Predictor
    xd, xw, xm = _grids(400, 3)
    rng = np.random.default_rng(5)
    mk = lambda x: gp.models.GPR((x, np.sin(x / 30.0) + 0.1 * rng.standard_normal(x.shape)),
                                 kernel=gp.kernels.SquaredExponential(lengthscales=25.0), noise_variance=1e-2)
    md, mw, mm = mk(xd), mk(xw), mk(xm)
    host = Predictor().predict_combined(0.5, 0.3, md, mw, mm, xd, xw, xm)
    models.set_output_device("cuda")
    try:
        dev = Predictor().predict_combined(0.5, 0.3, md, mw, mm, xd, xw, xm)
    finally:
        models.set_output_device("cpu")
    for h, d in zip(host, dev):
        assert d.is_cuda and np.array_equal(d.cpu().numpy(), np.asarray(h), equal_nan=True)
