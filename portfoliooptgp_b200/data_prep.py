"""Device-side data preparation: returns, z-scoring, column concat and the
rolling-window gather that produce the design matrix / the [B, N, D] batch directly in HBM.

Host mirror of Multi-Input_GPR/utils/data_handler.py's arithmetic (process_data :86-91,
normalize_and_reshape :160-179, concatenate_X :129-154) and of the window slicing in
Multi-Input_GPR/main.py:414-423, for price series that are already device tensors; CSV reading, date
filtering and the EODHD fetch stay on the host where the reference has them.  Every function takes
array-likes or CUDA tensors and returns fp64 CUDA tensors; all arithmetic runs in the kernels of
csrc/prep.cu through the C-ABI (gpb_prep_*).  No CPU fallback."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops

RETURN_KINDS = {"return": 0, "intraday_return": 1, "daily_log_return": 2}


def _engine_for(t: torch.Tensor):
    eng = ops.shared_engine(t.device.index)
    ops.sync_stream(eng)
    return eng


def _as_series(a, device=None) -> Tuple[torch.Tensor, bool]:
    t = ops.to_device(a, device)
    was_1d = t.ndim == 1
    if was_1d:
        t = t[:, None]
    if t.ndim != 2:
        raise ValueError("series must be [T] or [T, A]")
    return t.contiguous(), was_1d


def returns(close, open_=None, kind: str = "return", device=None) -> torch.Tensor:
    """The reference's three return columns (data_handler.py:86-91) for [T] or [T, A] price series.

    ``return``: ``close.pct_change()`` with the first row filled from the second row;
    ``intraday_return``: ``(close - open) / open``; ``daily_log_return``: ``log(close / close.shift(1))``
    with +-inf replaced by 0 and a NaN first row, as pandas leaves it."""
    if kind not in RETURN_KINDS:
        raise ValueError(f"kind must be one of {sorted(RETURN_KINDS)}")
    c, was_1d = _as_series(close, device)
    o = None
    if kind == "intraday_return":
        if open_ is None:
            raise ValueError("intraday_return needs the open series")
        o, _ = _as_series(open_, c.device.index)
        if o.shape != c.shape:
            raise ValueError("close and open must have the same shape")
    out = torch.empty_like(c)
    T, A = c.shape
    _engine_for(c).prep_returns(c.data_ptr(), None if o is None else o.data_ptr(), T, A, RETURN_KINDS[kind],
                                out.data_ptr())
    return out[:, 0] if was_1d else out


def zscore(x, ddof: int = 1, device=None, out: Optional[torch.Tensor] = None):
    """``(x - mean) / std`` per column with pandas' ddof=1 std (data_handler.py:160-169).
    Returns ``(z, mean, std)``; ``out`` may be a column block (a view with row stride > A) of a
    wider design matrix, which fuses concatenate_X."""
    t, was_1d = _as_series(x, device)
    T, A = t.shape
    mean = torch.empty(A, dtype=torch.float64, device=t.device)
    std = torch.empty(A, dtype=torch.float64, device=t.device)
    if out is None:
        out = torch.empty_like(t)
    else:
        if out.shape != t.shape or out.dtype != torch.float64 or out.device != t.device:
            raise ValueError("out must be an fp64 tensor of x's shape on x's device")
        if A > 1 and out.stride(1) != 1:
            raise ValueError("out must have unit column stride")
    ldo = out.stride(0) if T > 1 else max(A, out.stride(0))
    _engine_for(t).prep_zscore(t.data_ptr(), T, A, ddof, out.data_ptr(), ldo, mean.data_ptr(), std.data_ptr())
    if was_1d:
        return out[:, 0], mean[0], std[0]
    return out, mean, std


def normalize_and_reshape(y, x, device=None):
    """data_handler.py:160-179 for one (y_column, x_column) pair: returns
    ``X [T,1], Y [T,1], (y_mean, y_std), (x_mean, x_std)`` (the reference also passes the date column
    through; it never reaches the model)."""
    Y, ym, ys = zscore(y, device=device)
    X, xm, xs = zscore(x, device=Y.device.index)
    return X.reshape(-1, 1), Y.reshape(-1, 1), (ym, ys), (xm, xs)


def concatenate_X(X: Sequence) -> torch.Tensor:
    """data_handler.py:129-154: column concat of same-shaped [T,1] (or [T]) series into [T, len(X)].
    Same errors as the reference for a non-sequence, an empty list or mismatching shapes."""
    if not isinstance(X, (list, tuple)):
        raise ValueError("Input X should be a list or tuple of tensors")
    if len(X) < 1:
        raise ValueError("Input X should contain at least one tensor array")
    ts = [ops.to_device(x) for x in X]
    if not all(t.shape == ts[0].shape for t in ts):
        raise ValueError("All input tensors should have the same shape")
    return torch.cat([t.reshape(-1, 1) for t in ts], dim=1).contiguous()


def design_matrix(columns: Sequence, ddof: int = 1, device=None):
    """z-score every [T] / [T, A_k] block in ``columns`` straight into its slot of one [T, D] design
    matrix (normalize_and_reshape + concatenate_X without the intermediate copies).
    Returns ``(X, mean [D], std [D])``."""
    blocks = [_as_series(c, device)[0] for c in columns]
    if not blocks:
        raise ValueError("design_matrix needs at least one column block")
    T = blocks[0].shape[0]
    if any(b.shape[0] != T for b in blocks):
        raise ValueError("All column blocks must have the same number of rows")
    D = sum(b.shape[1] for b in blocks)
    dev = blocks[0].device
    X = torch.empty((T, D), dtype=torch.float64, device=dev)
    mean = torch.empty(D, dtype=torch.float64, device=dev)
    std = torch.empty(D, dtype=torch.float64, device=dev)
    eng = _engine_for(X)
    c0 = 0
    for b in blocks:
        A = b.shape[1]
        eng.prep_zscore(b.data_ptr(), T, A, ddof, X[:, c0:].data_ptr(), D, mean[c0:].data_ptr(), std[c0:].data_ptr())
        c0 += A
    return X, mean, std


def rolling_windows(features, y=None, window: int = 128, stride: int = 1, device=None):
    """Cut stride-``stride`` windows of length ``window`` from ``features`` [T, D] or [S, T, D] (and
    ``y`` [T] / [S, T]) into the batch ``X [B, window, D]``, ``Y [B, window, 1]`` with
    B = S * ((T - window) // stride + 1), series-major: the layout BatchedGPR / gpb_batched_lml_grad
    take (C3: 20 assets x 64 windows of 128 days)."""
    f = ops.to_device(features, device)
    if f.ndim == 2:
        f = f[None]
    if f.ndim != 3:
        raise ValueError("features must be [T, D] or [S, T, D]")
    S, T, D = f.shape
    yt = None
    if y is not None:
        yt = ops.to_device(y, f.device.index)
        yt = yt.reshape(S, T)
    if window < 1 or stride < 1:
        raise ValueError("window and stride must be >= 1")
    W = (T - window) // stride + 1 if T >= window else 0
    X = torch.empty((S * W, window, D), dtype=torch.float64, device=f.device)
    Y = None if yt is None else torch.empty((S * W, window, 1), dtype=torch.float64, device=f.device)
    if S * W > 0:
        _engine_for(f).prep_windows(f.data_ptr(), None if yt is None else yt.data_ptr(), S, T, D, window, stride,
                                    X.data_ptr(), None if Y is None else Y.data_ptr())
    return (X, Y) if y is not None else X


def expanding_windows(features, y, first: int, count: int, device=None):
    """The reference's rolling re-fit as ONE ragged batch (Multi-Input_GPR/main.py:414-423): test day k
    (k = 0 .. count-1) fits a fresh GPR on ``X_full[:first + k]``.  Returns ``(X [count, Nmax, D],
    Y [count, Nmax, 1], nrows [count], Xnew [count, 1, D])`` with Nmax = first + count - 1: every GP sees the
    same leading rows, ``nrows[k] = first + k`` says how many it uses, and ``Xnew[k]`` is the next row -- the
    one whose prediction the reference keeps (``predict_f(X_full[:i+1])[-1]``, main.py:434,454).  Feed
    ``BatchedGPR(X, Y, kernel, nrows=nrows)`` and ``predict_f(Xnew)``.  ``features`` needs first + count rows."""
    f = ops.to_device(features, device)
    if f.ndim != 2:
        raise ValueError("features must be [T, D]")
    yt = ops.to_device(y, f.device.index).reshape(-1)
    T, D = f.shape
    if first < 1 or count < 1 or first + count > T or yt.shape[0] != T:
        raise ValueError("expanding_windows needs 1 <= first, 1 <= count and first + count <= T rows of features and y")
    nmax = first + count - 1
    X = f[:nmax].unsqueeze(0).expand(count, nmax, D).contiguous()
    Y = yt[:nmax].reshape(1, nmax, 1).expand(count, nmax, 1).contiguous()
    nrows = np.arange(first, first + count, dtype=np.int32)
    Xnew = torch.stack([f[first + k] for k in range(count)])[:, None, :].contiguous()
    return X, Y, nrows, Xnew
