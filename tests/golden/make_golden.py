"""Generate tests/golden/*.npz: oracle outputs on fixed inputs for the reference's own call pattern.

GPflow 2.9.1 cannot be imported here (un-vendored, not installable), so these are outputs of the
CPU oracle (oracle/gpflow_oracle.py, GPflow-faithful Gram-form distances) -- they pin the oracle
against drift and give the GPU tests fixed targets; they are NOT outputs of GPflow itself
("parity unpinned", see the oracle header).

Inputs: the reference's real AAPL daily / weekly / monthly closes
(GPR/Stocks/AAPL_EOD/AAPL_us_{d,w,m}.csv, 89 / 19 / 5 rows) preprocessed exactly as
GPR/data_handler.py:26-66 does (day index from 2024-02-01, pct_change with the first value
back-filled, z-score with pandas' ddof=1 std), the 8 kernel candidates of GPR/main.py:105-114 at
GPflow's default hyper-parameters, sigma^2 = 1e-5 (GPR/model_trainer.py:16) and 1e-2; plus a
multi-input window (N=67, D=7, Exponential*Exponential, Multi-Input_GPR/main.py:126-135,422) and an
SVGP case (test_scripts/SVGP.py:515-521 shape: M=20 inducing points on a line).

Run from the repo root in the build container:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gpflow_oracle as O  # noqa: E402

REF = "/root/reference"


def load_aapl(period):
    df = pd.read_csv(f"{REF}/GPR/Stocks/AAPL_EOD/AAPL_us_{period}.csv")
    df["date"] = pd.to_datetime(df["date"])
    day = (df["date"] - pd.Timestamp("2024-02-01")).dt.days.values.astype(np.float64)
    ret = df["close"].pct_change()
    ret = ret.fillna(ret.iloc[1])
    y = ((ret - ret.mean()) / ret.std()).values.astype(np.float64)
    return day[:, None], y[:, None]


def reference_kernels():
    L = O.Leaf
    return {
        "SE": L("se"),
        "Matern12": L("matern12"),
        "RQ": L("rq"),
        "Exponential": L("exponential"),
        "SE+Matern12": O.Sum([L("se"), L("matern12")]),
        "Exp+Periodic(SE)+Linear": O.Sum([L("exponential"), O.Periodic(L("se"), 1.0), L("linear")]),
        "Exp+Periodic(SE)": O.Sum([L("exponential"), O.Periodic(L("se"), 1.0)]),
        "SE*Matern12": O.Product([L("se"), L("matern12")]),
    }


def main():
    out = {}
    kernels = reference_kernels()
    out["kernel_names"] = np.array(list(kernels))
    for period in ("d", "w", "m"):
        X, Y = load_aapl(period)
        Xs = np.concatenate([X, X[-1:] + np.arange(1, 31)[:, None]], axis=0)   # train + 30 future days (GPR/main.py:68-71)
        out[f"aapl_{period}_X"], out[f"aapl_{period}_Y"], out[f"aapl_{period}_Xs"] = X, Y, Xs
        for name, k in kernels.items():
            for tag, s2 in (("1e-5", 1e-5), ("1e-2", 1e-2)):
                lml, g, gn = O.gpr_lml_and_grad(k, X, Y, s2)
                mean, var = O.gpr_predict_f(k, X, Y, s2, Xs)
                Kn = O.K(k, X) + s2 * np.eye(len(X))
                key = f"aapl_{period}|{name}|{tag}"
                out[key + "|lml"] = np.array(lml)
                out[key + "|grad"] = np.concatenate([g, [gn]])
                out[key + "|mean"], out[key + "|var"] = mean, var
                out[key + "|cond"] = np.array(np.linalg.cond(Kn))
    # multi-input rolling window
    rng = np.random.default_rng(7)
    N, D = 67, 7
    f = rng.normal(0, 0.01, (N + 1, 1))
    r = rng.uniform(0.5, 1.5, (1, D)) * f + rng.normal(0, 0.01, (N + 1, D))
    z = lambda a: (a - a.mean(0)) / a.std(0)
    Xm = np.concatenate([z(r[:, 1:]), z(np.arange(N + 1, dtype=np.float64)[:, None])], axis=1)
    Ym = z(r[:, :1])
    km = O.Product([O.Leaf("exponential", active_dims=slice(0, D - 1)), O.Leaf("exponential", active_dims=slice(D - 1, D))])
    out["multi_X"], out["multi_Y"] = Xm, Ym
    lml, g, gn = O.gpr_lml_and_grad(km, Xm[:N], Ym[:N], 1e-3)
    mean, var = O.gpr_predict_f(km, Xm[:N], Ym[:N], 1e-3, Xm)
    out["multi|lml"], out["multi|grad"], out["multi|mean"], out["multi|var"] = np.array(lml), np.concatenate([g, [gn]]), mean, var
    # SVGP
    X, Y = load_aapl("d")
    M = 20
    Z = np.linspace(0, X.max(), M)[:, None]
    qmu = 0.1 * np.sin(np.arange(M))[:, None]
    qs = (0.8 * np.eye(M) + 0.02 * np.tril(np.cos(np.add.outer(np.arange(M), 2.0 * np.arange(M)))))[None]
    ks = O.Leaf("se", 1.0, 10.0)
    out["svgp_Z"], out["svgp_qmu"], out["svgp_qsqrt"] = Z, qmu, qs
    out["svgp|elbo"] = np.array(O.svgp_elbo(ks, Z, qmu, qs, 1e-4 * 100, X, Y, num_data=len(X)))
    m, v = O.svgp_predict_f(ks, Z, qmu, qs, X)
    out["svgp|mean"], out["svgp|var"] = m, v
    path = os.path.join(ROOT, "tests", "golden", "reference_call_pattern.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
