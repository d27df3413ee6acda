"""Prediction post-processing of the reference's Predictor (GPR/predictor.py:10-51):
linear-interpolation upsampling of the weekly / monthly predictions onto the daily grid and the
alpha / beta blend of the three time frames, plus the SLSQP solve for the blend weights
(GPR/optimizer.py:5-28).

Two paths with the same results, bit for bit: when the predictions are CUDA tensors (predict_f with
``models.set_output_device("cuda")``) upsampling and blend run on the device (csrc/prep.cu
``gpb_post_upsample`` / ``gpb_post_blend``; one launch per step, nothing crosses to the host); host arrays
(the reference's few hundred values) take the NumPy path below.  The two-variable SLSQP stays on the host."""
from __future__ import annotations

import numpy as np


def _is_cuda(a) -> bool:
    return hasattr(a, "is_cuda") and bool(a.is_cuda)


def _dev_vec(a, device):
    import torch
    t = a if hasattr(a, "is_cuda") else torch.as_tensor(np.asarray(a, dtype=np.float64))
    return t.detach().to(device=device, dtype=torch.float64).reshape(-1).contiguous()


def upsample_predictions_device(X_daily, X, predictions, period: str = "d"):
    """Device form of ``upsample_predictions``: ``predictions`` is one [Ns, 1] CUDA tensor or a sequence
    of Q of them (all four outputs of predict_single share one search).  Grids ascending and unique."""
    import torch
    from . import ops
    single = hasattr(predictions, "is_cuda")
    preds = [predictions] if single else list(predictions)
    if period not in ("w", "m"):
        return predictions
    device = next(p.device for p in preds if _is_cuda(p))
    eng = ops.shared_engine(device.index)
    ops.sync_stream(eng)
    xd, xs = _dev_vec(X_daily, device), _dev_vec(X, device)
    P = torch.stack([_dev_vec(p, device) for p in preds])                # [Q, Ns]
    out = torch.empty((len(preds), xd.numel()), dtype=torch.float64, device=device)
    eng.post_upsample(xd.data_ptr(), xd.numel(), xs.data_ptr(), xs.numel(), P.data_ptr(), len(preds), out.data_ptr())
    cols = [out[q][:, None] for q in range(len(preds))]
    return cols[0] if single else tuple(cols)


def predict_combined_device(alpha, beta, daily, weekly, monthly, X_daily, X_weekly, X_monthly):
    """Device form of ``predict_combined``: two upsampling launches (weekly, monthly; four columns each)
    and one blend launch over all four columns."""
    import torch
    from . import ops
    device = next(t.device for t in daily if _is_cuda(t))
    wu = upsample_predictions_device(X_daily, X_weekly, tuple(weekly), period="w")
    mu = upsample_predictions_device(X_daily, X_monthly, tuple(monthly), period="m")
    D = torch.stack([_dev_vec(d, device) for d in daily])
    Wu = torch.stack([w.reshape(-1) for w in wu]).contiguous()
    Mu = torch.stack([m.reshape(-1) for m in mu]).contiguous()
    out = torch.empty_like(D)
    eng = ops.shared_engine(device.index)
    ops.sync_stream(eng)
    eng.post_blend(float(alpha), float(beta), D.data_ptr(), Wu.data_ptr(), Mu.data_ptr(), D.numel(), out.data_ptr())
    return tuple(out[q][:, None] for q in range(out.shape[0]))


def _np(a):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    elif hasattr(a, "numpy") and not isinstance(a, np.ndarray):
        a = a.numpy()
    return np.asarray(a, dtype=np.float64)


def upsample_predictions(X_daily, X, predictions, period: str = "d"):
    """GPR/predictor.py:35-51: ``pd.Series(pred, index=X).reindex(X_daily).interpolate('linear')``.
    pandas' 'linear' ignores the index and interpolates over POSITIONS in the daily grid; leading
    gaps stay NaN, trailing gaps repeat the last value."""
    if period not in ("w", "m"):
        return predictions
    if _is_cuda(predictions):
        return upsample_predictions_device(X_daily, X, predictions, period)
    xd = _np(X_daily).reshape(-1)
    x = _np(X).reshape(-1)
    p = _np(predictions).reshape(-1)
    lookup = {}
    for xi, pi in zip(x, p):
        lookup.setdefault(xi, pi)   # reindex on a unique index; duplicates are not produced by the reference
    vals = np.array([lookup.get(v, np.nan) for v in xd], dtype=np.float64)
    pos = np.arange(len(xd), dtype=np.float64)
    ok = ~np.isnan(vals)
    out = vals.copy()
    if ok.any():
        first, last = np.argmax(ok), len(ok) - 1 - np.argmax(ok[::-1])
        inner = slice(first, last + 1)
        out[inner] = np.interp(pos[inner], pos[ok], vals[ok])
        out[last + 1:] = vals[last]
    return out.reshape(-1, 1)


def predict_combined(alpha, beta, daily, weekly, monthly, X_daily, X_weekly, X_monthly):
    """GPR/predictor.py:10-33.  ``daily`` / ``weekly`` / ``monthly`` are the 4-tuples
    (f_mean, f_var, y_mean, y_var) of Predictor.predict_single for each model."""
    if any(_is_cuda(t) for t in daily):
        return predict_combined_device(alpha, beta, daily, weekly, monthly, X_daily, X_weekly, X_monthly)
    out = []
    for d, w, m in zip(daily, weekly, monthly):
        wu = upsample_predictions(X_daily, X_weekly, w, period="w")
        mu = upsample_predictions(X_daily, X_monthly, m, period="m")
        out.append(alpha * _np(d) + beta * wu + (1 - alpha - beta) * mu)
    return tuple(out)


class Predictor:
    """Drop-in for GPR/predictor.py's class (same method names and argument order)."""

    def predict_single(self, model, X):
        f_mean, f_var = model.predict_f(X, full_cov=False)
        y_mean, y_var = model.predict_y(X)
        return f_mean, f_var, y_mean, y_var

    def predict_combined(self, alpha, beta, daily_model, weekly_model, monthly_model, X_daily, X_weekly, X_monthly):
        return predict_combined(alpha, beta, self.predict_single(daily_model, X_daily),
                                self.predict_single(weekly_model, X_weekly), self.predict_single(monthly_model, X_monthly),
                                X_daily, X_weekly, X_monthly)

    def upsample_predictions(self, X_daily, X, predictions, period="d"):
        return upsample_predictions(X_daily, X, predictions, period)


class Optimizer:
    """Blend-weight solve of GPR/optimizer.py:5-28: minimise  MSE(Y, a*daily + b*weekly + (1-a-b)*monthly)
    + lambda (|a| + |b|)  over 0 <= a, b <= 1, a + b <= 1 with SciPy's SLSQP from (0.33, 0.33).  A
    two-variable host problem on predictions that are already on the host; same attribute and
    method names as the reference class."""

    def __init__(self, lambda_=0.01):
        self.lambda_ = lambda_
        self.initial_weights = [0.33, 0.33]
        self.bounds = [(0, 1), (0, 1)]
        self.constraints = {"type": "ineq", "fun": lambda w: 1 - sum(w)}

    def loss_fn(self, weights, Y, f_mean_daily, f_mean_weekly, f_mean_monthly):
        a, b = weights
        resid = _np(Y) - (a * _np(f_mean_daily) + b * _np(f_mean_weekly) + (1 - a - b) * _np(f_mean_monthly))
        # sklearn's mean_squared_error on [N,1] columns: mean over rows, uniform average over outputs
        return float(np.mean(resid * resid)) + self.lambda_ * (abs(a) + abs(b))

    def optimize_weights(self, Y_tf, f_mean_daily, f_mean_weekly, f_mean_monthly):
        from scipy.optimize import minimize

        res = minimize(lambda w: self.loss_fn(w, Y_tf, f_mean_daily, f_mean_weekly, f_mean_monthly),
                       self.initial_weights, bounds=self.bounds, constraints=self.constraints, method="SLSQP")
        return res.x
