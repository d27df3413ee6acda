"""Shared test helpers: package kernel objects -> oracle descriptions, synthetic data."""
import numpy as np

import portfoliooptgp_b200 as gpflow
from oracle import gpflow_oracle as O

_KIND = {
    gpflow.kernels.SquaredExponential: "se",
    gpflow.kernels.RationalQuadratic: "rq",
    gpflow.kernels.Matern12: "matern12",
    gpflow.kernels.Exponential: "exponential",
    gpflow.kernels.Matern32: "matern32",
    gpflow.kernels.Matern52: "matern52",
}


def _val(p):
    v = p.numpy()
    return float(v) if v.ndim == 0 else v.copy()


def to_oracle(k):
    """Translate a package kernel tree to the oracle's plain-data tree (same parameter order)."""
    K = gpflow.kernels
    if isinstance(k, K.Sum):
        return O.Sum([to_oracle(c) for c in k.kernels])
    if isinstance(k, K.Product):
        return O.Product([to_oracle(c) for c in k.kernels])
    if isinstance(k, K.Periodic):
        return O.Periodic(to_oracle(k.base_kernel), _val(k.period))
    if isinstance(k, K.Linear):
        return O.Leaf("linear", variance=_val(k.variance), active_dims=k.active_dims)
    kind = _KIND[type(k)]
    return O.Leaf(kind, variance=_val(k.variance), lengthscales=_val(k.lengthscales),
                  alpha=_val(k.alpha) if kind == "rq" else 1.0, active_dims=k.active_dims)


def zscore(a):
    return (a - a.mean(0)) / a.std(0)


def factor_returns(rng, n, d):
    """Factor-model daily returns: r_tj = beta_j f_t + eps_tj."""
    f = rng.normal(0.0, 0.01, size=(n, 1))
    beta = rng.uniform(0.5, 1.5, size=(1, d))
    return beta * f + rng.normal(0.0, 0.01, size=(n, d))


def make_multi_input(seed, n, d):
    """X [n,d] = (d-1) z-scored feature-return columns + z-scored time; y = z-scored target returns."""
    rng = np.random.default_rng(seed)
    r = factor_returns(rng, n, d)
    t = np.arange(n, dtype=np.float64)[:, None]
    X = np.concatenate([zscore(r[:, 1:]), zscore(t)], axis=1) if d > 1 else zscore(t)
    y = zscore(r[:, :1])
    return np.ascontiguousarray(X), np.ascontiguousarray(y)


def make_c1(n=1000, seed=1):
    """SURVEY.md 8d C1 (BASELINE config 1): z-scored day index, z-scored synthetic daily returns with a
    21-day cycle; the same generator as bench.py::make_c1 and tests/golden/make_truth.py::make_c1."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64)[:, None]
    X = (t - t.mean()) / t.std()
    r = rng.normal(0, 0.01, size=(n, 1)) + 0.004 * np.sin(2 * np.pi * t / 21.0)
    Y = (r - r.mean()) / r.std()
    return np.ascontiguousarray(X), np.ascontiguousarray(Y)


def record_parity(name, payload):
    """Append one measured-parity record to gpurun_out/parity_measured.jsonl (travels back from the GPU
    box; the numbers quoted in DESIGN.md section 10 come from this file)."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_measured.jsonl"), "a") as f:
            f.write(json.dumps({"case": name, **payload}) + "\n")
    except OSError:
        pass


def kernel_zoo(D):
    """Kernel expressions covering the reference call sites and every leaf/group kind."""
    K = gpflow.kernels
    last = [D - 1]
    zoo = {
        "se": K.SquaredExponential(variance=1.3, lengthscales=0.9),
        "matern12": K.Matern12(variance=0.8, lengthscales=1.4),
        "rq": K.RationalQuadratic(variance=1.1, lengthscales=1.2, alpha=0.7),
        "exponential": K.Exponential(variance=0.9, lengthscales=1.1),
        "se+matern12": K.SquaredExponential(variance=1.2, lengthscales=0.8) + K.Matern12(variance=0.5, lengthscales=2.0),
        "exp+periodic(se)+linear": K.Exponential(variance=0.7, lengthscales=1.3)
        + K.Periodic(K.SquaredExponential(variance=0.6, lengthscales=1.1, active_dims=last), period=1.7)
        + K.Linear(variance=0.3),
        "se*matern12": K.SquaredExponential(variance=1.2, lengthscales=0.8) * K.Matern12(variance=0.5, lengthscales=2.0),
        "se+matern52+linear": K.SquaredExponential(variance=1.0, lengthscales=1.5) + K.Matern52(variance=0.7, lengthscales=2.5)
        + K.Linear(variance=0.2),
        "matern32": K.Matern32(variance=0.9, lengthscales=1.7),
        "periodic(matern32)": K.Periodic(K.Matern32(variance=0.9, lengthscales=1.3, active_dims=last), period=2.1),
    }
    if D >= 2:
        zoo["exp[0:D-1]*exp[D-1]"] = (K.Exponential(variance=1.1, lengthscales=1.6, active_dims=slice(0, D - 1))
                                       * K.Exponential(variance=0.8, lengthscales=0.9, active_dims=slice(D - 1, D)))
        zoo["ard_se"] = K.SquaredExponential(variance=1.1, lengthscales=np.linspace(0.8, 2.0, D))
        zoo["(se+lin)*exp"] = ((K.SquaredExponential(variance=0.9, lengthscales=1.2, active_dims=slice(0, D - 1))
                                + K.Linear(variance=0.4, active_dims=slice(0, D - 1)))
                               * K.Exponential(variance=1.0, lengthscales=2.0, active_dims=[D - 1]))
    return zoo
