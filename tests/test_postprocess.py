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
