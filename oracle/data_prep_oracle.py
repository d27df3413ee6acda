"""CPU oracle for the device-side data preparation (csrc/prep.cu).  TEST INFRASTRUCTURE ONLY: imported
by tests/ alone, never by the product path.

It issues the same pandas calls as the reference (Multi-Input_GPR/utils/data_handler.py), so for
these few lines the oracle IS the reference arithmetic; the reference class itself cannot be
imported here (it pulls in tensorflow, dotenv and OrdinalEntroPy at module import)."""
import numpy as np
import pandas as pd


def returns(close, open_=None, kind="return"):
    """data_handler.py:86-91, per column of a [T] / [T, A] array."""
    close = np.asarray(close, dtype=np.float64)
    cols = close.reshape(len(close), -1)
    out = np.empty_like(cols)
    for a in range(cols.shape[1]):
        df = pd.DataFrame({"close": cols[:, a]})
        if kind == "return":
            df["return"] = df["close"].pct_change()                 # :86
            first_return = df["return"].iloc[1]                      # :87
            df.fillna({"return": first_return}, inplace=True)        # :88
            out[:, a] = df["return"].values
        elif kind == "intraday_return":
            df["open"] = np.asarray(open_, dtype=np.float64).reshape(len(close), -1)[:, a]
            out[:, a] = ((df["close"] - df["open"]) / df["open"]).values   # :89
        elif kind == "daily_log_return":
            r = np.log(df["close"] / df["close"].shift(1))          # :90
            out[:, a] = r.replace([np.inf, -np.inf], 0).values       # :91
        else:
            raise ValueError(kind)
    return out.reshape(close.shape)


def zscore(x):
    """data_handler.py:160-169: pandas mean / std (ddof = 1) per column."""
    x = np.asarray(x, dtype=np.float64)
    cols = x.reshape(len(x), -1)
    df = pd.DataFrame(cols)
    mean, std = df.mean().values, df.std().values
    return ((cols - mean) / std).reshape(x.shape), mean, std


def concatenate_X(X):
    """data_handler.py:129-154."""
    return np.concatenate([np.asarray(x).reshape(-1, 1) for x in X], axis=1)


def rolling_windows(features, y, window, stride):
    """Stride-s windows of length `window` per series, series-major (C3 layout)."""
    f = np.asarray(features, dtype=np.float64)
    if f.ndim == 2:
        f = f[None]
    S, T, D = f.shape
    yy = None if y is None else np.asarray(y, dtype=np.float64).reshape(S, T)
    Xs, Ys = [], []
    for s in range(S):
        for w0 in range(0, T - window + 1, stride):
            Xs.append(f[s, w0:w0 + window])
            if yy is not None:
                Ys.append(yy[s, w0:w0 + window, None])
    X = np.stack(Xs) if Xs else np.empty((0, window, D))
    if yy is None:
        return X
    return X, (np.stack(Ys) if Ys else np.empty((0, window, 1)))
