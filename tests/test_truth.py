"""Extended-precision truth (tests/golden/extended_precision_truth.npz, made by
tests/golden/make_truth.py: mpmath at 50 digits for N = 89 / 300, numpy.longdouble for N = 1000).

VERDICT r01 missing #1: at the reference's own sigma^2 = 1e-5 (GPR/model_trainer.py:16) the fp64
oracle was the only "truth", so nobody knew which side of a GPU-vs-oracle gap was closer.  Here both
fp64 implementations are measured against a truth that is more accurate than either:

  CPU part : |oracle - truth| for every case (bounds = what LAPACK fp64 achieves here, measured).
  GPU part : |CUDA - truth| under the SAME bounds -- the engine (explicit inverses of the diagonal
             blocks instead of TRSM, direct-difference distances, rsqrt pivots) has to be as close to
             the truth as the LAPACK oracle is, case by case -- and under the north-star bars
             (1e-9 values, 1e-7 gradients) wherever fp64 LAPACK itself meets them.

Gradient errors are measured relative to the largest gradient component (the scale L-BFGS sees)."""
import os

import numpy as np
import pytest

from oracle import gpflow_oracle as O
from tests.helpers import record_parity

T = np.load(os.path.join(os.path.dirname(__file__), "golden", "extended_precision_truth.npz"))

NAMES = ["SE", "Matern12", "RQ", "Exponential", "SE+Matern12", "Exp+Periodic(SE)+Linear", "Exp+Periodic(SE)", "SE*Matern12"]
CASES = [(f"aapl_d|{n}", n, "mp") for n in NAMES] + [("c1_300|SE+Periodic(SE)", "SE+Periodic(SE)", "mp"),
                                                       ("c1_1000|SE+Periodic(SE)", "SE+Periodic(SE)", "ld")]

# Bars per noise level: (LML rel, gradient rel-to-max, mean rel-to-max, variance abs / max(var, 1e-3)).
# 1e-2: the north-star bars.  1e-5: cond(K + s2 I) reaches 1.4e8 on the C1 axis; fp64 LAPACK (the oracle,
# either distance form) is measured below at <= 1.6e-10 / 3.2e-10 / 1.2e-8 / 2e-12 -- the predictive MEAN
# cannot meet 1e-9 in fp64 at this conditioning (alpha = K^-1 y carries cond * eps), everything else does.
# The CUDA path is held to the same bars (measured at N = 1000, sigma^2 = 1e-5: 1.7e-12 / 3.7e-12 / 3.8e-9 /
# 1.3e-10, i.e. closer to the truth than LAPACK on LML, gradient and mean; profiles/r02_parity_measured.jsonl).
BARS = {"1e-2": (1e-9, 1e-7, 1e-9, 1e-9), "1e-5": (1e-9, 1e-7, 5e-8, 1e-9)}
GPU_BARS = BARS


def oracle_kernel(name):
    L = O.Leaf
    return {
        "SE": L("se"), "Matern12": L("matern12"), "RQ": L("rq"), "Exponential": L("exponential"),
        "SE+Matern12": O.Sum([L("se"), L("matern12")]),
        "Exp+Periodic(SE)+Linear": O.Sum([L("exponential"), O.Periodic(L("se"), 1.0), L("linear")]),
        "Exp+Periodic(SE)": O.Sum([L("exponential"), O.Periodic(L("se"), 1.0)]),
        "SE*Matern12": O.Product([L("se"), L("matern12")]),
        "SE+Periodic(SE)": O.Sum([L("se"), O.Periodic(L("se"), 1.0)]),
    }[name]


def gp_kernel(gpflow, name):
    K = gpflow.kernels
    return {
        "SE": lambda: K.SquaredExponential(), "Matern12": lambda: K.Matern12(), "RQ": lambda: K.RationalQuadratic(),
        "Exponential": lambda: K.Exponential(), "SE+Matern12": lambda: K.SquaredExponential() + K.Matern12(),
        "Exp+Periodic(SE)+Linear": lambda: K.Exponential() + K.Periodic(K.SquaredExponential()) + K.Linear(),
        "Exp+Periodic(SE)": lambda: K.Exponential() + K.Periodic(K.SquaredExponential()),
        "SE*Matern12": lambda: K.SquaredExponential() * K.Matern12(),
        "SE+Periodic(SE)": lambda: K.SquaredExponential() + K.Periodic(K.SquaredExponential()),
    }[name]()


def errors(key, tag, be, lml, grad, mean, var):
    k = f"{key}|{tag}|{be}"
    l0, g0, m0, v0 = float(T[k + "|lml"]), T[k + "|grad"], T[k + "|mean"], T[k + "|var"]
    return {"lml_rel": abs(lml - l0) / abs(l0),
            "grad_rel_to_max": float(np.max(np.abs(np.asarray(grad) - g0)) / np.max(np.abs(g0))),
            "mean_rel_to_max": float(np.max(np.abs(mean - m0)) / np.max(np.abs(m0))),
            "var_abs_scaled": float(np.max(np.abs(var - v0)) / max(1e-3, float(np.max(np.abs(v0)))))}


def check(err, tag, what, bars_by_tag=BARS):
    bars = bars_by_tag[tag]
    for (name, val), bar in zip(err.items(), bars):
        assert val <= bar, (what, tag, name, val, bar)


@pytest.mark.parametrize("form", ["gram", "direct"])
def test_oracle_against_truth(form):
    """The fp64 LAPACK oracle (GPflow's Gram-form distances and the direct form the CUDA kernels use) against
    the extended-precision truth, every case, both noise levels."""
    O.set_distance_form(form)
    try:
        for key, name, be in CASES:
            X, Y, Xs = T[key + "|X"], T[key + "|Y"], T[key + "|Xs"]
            k = oracle_kernel(name)
            for tag, s2 in (("1e-5", 1e-5), ("1e-2", 1e-2)):
                lml, g, gn = O.gpr_lml_and_grad(k, X, Y, s2)
                mean, var = O.gpr_predict_f(k, X, Y, s2, Xs)
                check(errors(key, tag, be, lml, np.concatenate([g, [gn]]), mean, var), tag, (key, "oracle", form))
    finally:
        O.set_distance_form("gram")


def test_longdouble_backend_agrees_with_mpmath():
    """The two extended-precision back ends agree far below fp64 resolution wherever both ran (so the
    longdouble-only N = 1000 truth can be trusted)."""
    for key, name, be in CASES:
        if be != "mp":
            continue
        for tag in ("1e-5", "1e-2"):
            a, b = f"{key}|{tag}|mp", f"{key}|{tag}|ld"
            assert abs(float(T[a + "|lml"]) - float(T[b + "|lml"])) <= 1e-13 * abs(float(T[a + "|lml"]))
            assert np.max(np.abs(T[a + "|grad"] - T[b + "|grad"])) <= 1e-12 * np.max(np.abs(T[a + "|grad"]))
            assert np.max(np.abs(T[a + "|mean"] - T[b + "|mean"])) <= 1e-11 * np.max(np.abs(T[a + "|mean"]))
            assert np.max(np.abs(T[a + "|var"] - T[b + "|var"])) <= 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("key,name,be", CASES)
def test_gpu_against_truth(gp, key, name, be):
    """GPR/model_trainer.py:15-20 (noise assigned and frozen, objective + gradient, predict_f) on the CUDA
    path vs the extended-precision truth; the oracle's own error on the same case is recorded next to it."""
    X, Y, Xs = T[key + "|X"], T[key + "|Y"], T[key + "|Xs"]
    ko = oracle_kernel(name)
    for tag, s2 in (("1e-5", 1e-5), ("1e-2", 1e-2)):
        model = gp.models.GPR(data=(X, Y), kernel=gp_kernel(gp, name))
        model.likelihood.variance.assign(s2)
        gp.set_trainable(model.likelihood.variance, False)
        lml, g, gn = model.lml_and_constrained_grads()
        mean, var = model.predict_f(Xs, full_cov=False)
        e_gpu = errors(key, tag, be, lml, np.concatenate([g, [gn]]), mean.numpy(), var.numpy())
        lo, go, gno = O.gpr_lml_and_grad(ko, X, Y, s2)
        mo, vo = O.gpr_predict_f(ko, X, Y, s2, Xs)
        e_cpu = errors(key, tag, be, lo, np.concatenate([go, [gno]]), mo, vo)
        record_parity("truth|" + key + "|" + tag, {"gpu": e_gpu, "oracle_fp64_lapack": e_cpu, "n": int(len(X))})
        check(e_gpu, tag, (key, "gpu"), GPU_BARS)
