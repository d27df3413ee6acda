"""Generate tests/golden/extended_precision_truth.npz: exact-GP log marginal likelihood, gradient and
predictive moments computed in EXTENDED PRECISION, so that the fp64 CPU oracle and the fp64 CUDA
path can both be measured against something more accurate than either (missing #1:
at the reference's own sigma^2 = 1e-5, GPR/model_trainer.py:16, nobody knew which side is closer).

Two independent arithmetic back ends share one kernel / GP formulation (this file is a THIRD
statement of the maths, written against the published GPflow 2.9.1 formulas -- SURVEY.md G3-G11 --
not against oracle/gpflow_oracle.py):

  * mpmath, 50 significant digits: N = 89 (the reference's AAPL daily series, all 8 candidate
    kernels of GPR/main.py:105-114) and N = 300 (the C1 series cut to 300 points);
  * numpy.longdouble (x87, 64-bit significand, eps = 1.1e-19): the same cases (cross-check of the
    back end against mpmath) and C1 at its full N = 1000.

Both sigma^2 = 1e-5 (reference-faithful) and 1e-2.  Stored values are the extended-precision results
rounded once to fp64.  Run from the repo root in the build container (reads the reference's CSVs):

    python tests/golden/make_truth.py            # ~10 minutes, mostly the N = 300 mpmath cases
"""
import os
import sys
import time

import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

LD = np.longdouble


# ---- data ----------------------------------------------------------------------------------------------
def make_c1(n=1000, seed=1):
    """z-scored day index, z-scored synthetic daily returns with a 21-day cycle
    (the same generator as bench.py / tests.helpers.make_c1)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64)[:, None]
    X = (t - t.mean()) / t.std()
    r = rng.normal(0, 0.01, size=(n, 1)) + 0.004 * np.sin(2 * np.pi * t / 21.0)
    Y = (r - r.mean()) / r.std()
    return X, Y


# ---- kernels: elementwise on the difference d = x - x' and the pair (x, x'), generic arithmetic -----------
class MathMP:
    exp, sin, cos, sqrt, log = mp.exp, mp.sin, mp.cos, mp.sqrt, mp.log
    pi = None  # set after mp.dps

    @staticmethod
    def absf(v):
        return abs(v)


class MathLD:
    exp, sin, cos, sqrt, log = np.exp, np.sin, np.cos, np.sqrt, np.log
    pi = None  # set in gp_longdouble

    @staticmethod
    def absf(v):
        return np.abs(v)


def leaf(kind, M, d, x, x2, th):
    """(k, [dk/dtheta...]) for one leaf at constrained parameters th (GPflow order: alpha, lengthscales,
    variance; Periodic: base lengthscales, base variance, period; Linear: variance)."""
    if kind == "linear":
        (v,) = th
        return v * x * x2, [x * x2]
    if kind == "periodic_se":
        l, v, p = th
        arg = M.pi * d / p
        s = M.sin(arg) / l
        k = v * M.exp(-s * s / 2)
        return k, [k * s * s / l, k / v, k * s * M.cos(arg) * M.pi * d / (p * p * l)]
    if kind == "se":
        l, v = th
        r2 = d * d / (l * l)
        k = v * M.exp(-r2 / 2)
        return k, [k * r2 / l, k / v]
    if kind == "rq":
        a, l, v = th
        r2 = d * d / (l * l)
        base = 1 + r2 / (2 * a)
        k = v * M.exp(-a * M.log(base))
        return k, [k * (-M.log(base) + r2 / (2 * a) / base), v * M.exp(-(a + 1) * M.log(base)) * r2 / l, k / v]
    if kind in ("matern12", "exponential"):
        l, v = th
        c = 1 if kind == "matern12" else 2
        r = M.absf(d) / l
        k = v * M.exp(-r / c)
        return k, [k * r / (c * l), k / v]
    raise ValueError(kind)


N_PARAMS = {"linear": 1, "periodic_se": 3, "se": 2, "rq": 3, "matern12": 2, "exponential": 2}

# GPR/main.py:105-114, in that order; ("sum"|"prod", [leaf kinds])
KERNELS = {
    "SE": ("sum", ["se"]),
    "Matern12": ("sum", ["matern12"]),
    "RQ": ("sum", ["rq"]),
    "Exponential": ("sum", ["exponential"]),
    "SE+Matern12": ("sum", ["se", "matern12"]),
    "Exp+Periodic(SE)+Linear": ("sum", ["exponential", "periodic_se", "linear"]),
    "Exp+Periodic(SE)": ("sum", ["exponential", "periodic_se"]),
    "SE*Matern12": ("prod", ["se", "matern12"]),
    "SE+Periodic(SE)": ("sum", ["se", "periodic_se"]),   # C1 (BASELINE config 1)
}


def kernel_eval(spec, M, d, x, x2, one):
    """All parameters at GPflow's default 1.0.  Returns (k, [dk/dtheta_p])."""
    op, kinds = spec
    parts = [leaf(kd, M, d, x, x2, [one] * N_PARAMS[kd]) for kd in kinds]
    if op == "sum":
        k = parts[0][0]
        for q in parts[1:]:
            k = k + q[0]
        return k, [g for q in parts for g in q[1]]
    (k1, g1), (k2, g2) = parts
    return k1 * k2, [g * k2 for g in g1] + [k1 * g for g in g2]


# ---- numpy.longdouble back end --------------------------------------------------------------------------
def gp_longdouble(spec, X, Y, Xs, s2):
    MathLD.pi = LD(mp.nstr(mp.pi, 25))
    x = X[:, 0].astype(LD)
    y = Y[:, 0].astype(LD)
    n = len(x)
    d = x[:, None] - x[None, :]
    Kmat, dKs = kernel_eval(spec, MathLD, d, x[:, None], x[None, :], LD(1))
    A = Kmat + LD(s2) * np.eye(n, dtype=LD)
    L = np.zeros((n, n), dtype=LD)
    A = A.copy()
    for j in range(n):                       # right-looking Cholesky, vectorised rank-1 updates
        L[j, j] = np.sqrt(A[j, j])
        L[j + 1:, j] = A[j + 1:, j] / L[j, j]
        A[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], L[j + 1:, j])
    Li = np.zeros((n, n), dtype=LD)          # L^-1 by forward substitution, row by row
    for i in range(n):
        Li[i, i] = 1 / L[i, i]
        if i:
            Li[i, :i] = -(L[i, :i] @ Li[:i, :i]) / L[i, i]
    a = Li @ y
    alpha = Li.T @ a
    lml = -(a @ a) / 2 - LD(n) / 2 * np.log(2 * MathLD.pi) - np.sum(np.log(np.diag(L)))
    Kinv = Li.T @ Li
    Wm = np.outer(alpha, alpha) - Kinv
    grad = [np.sum(Wm * dK) / 2 for dK in dKs] + [np.trace(Wm) / 2]
    xs = Xs[:, 0].astype(LD)
    Ks, _ = kernel_eval(spec, MathLD, x[:, None] - xs[None, :], x[:, None], xs[None, :], LD(1))
    kss, _ = kernel_eval(spec, MathLD, np.zeros_like(xs), xs, xs, LD(1))
    Am = Li @ Ks
    mean = Am.T @ a
    var = kss - np.sum(Am * Am, axis=0)
    return np.float64(lml), np.array(grad, dtype=np.float64), mean.astype(np.float64), var.astype(np.float64)


# ---- mpmath back end ------------------------------------------------------------------------------------
def gp_mpmath(spec, X, Y, Xs, s2, want_grad=True):
    mp.mp.dps = 50
    MathMP.pi = mp.pi
    x = [mp.mpf(float(v)) for v in X[:, 0]]
    y = [mp.mpf(float(v)) for v in Y[:, 0]]
    n = len(x)
    one = mp.mpf(1)
    P = sum(N_PARAMS[k] for k in spec[1])
    Kl = [[None] * (i + 1) for i in range(n)]
    dK = [[[None] * (i + 1) for i in range(n)] for _ in range(P)]
    for i in range(n):
        for j in range(i + 1):
            k, gs = kernel_eval(spec, MathMP, x[i] - x[j], x[i], x[j], one)
            Kl[i][j] = k
            for p in range(P):
                dK[p][i][j] = gs[p]
        Kl[i][i] += mp.mpf(s2)
    L = [[mp.mpf(0)] * n for _ in range(n)]
    for i in range(n):                       # Cholesky-Banachiewicz with fdot row products
        Li_ = L[i]
        for j in range(i + 1):
            s = Kl[i][j] - mp.fdot(Li_[:j], L[j][:j])
            Li_[j] = mp.sqrt(s) if i == j else s / L[j][j]
    Linv = [[mp.mpf(0)] * n for _ in range(n)]
    LinvT_cols = [[] for _ in range(n)]      # column c of Linv as a growing list (rows c..i-1), for fdot
    for i in range(n):
        Linv[i][i] = 1 / L[i][i]
        for c in range(i):
            # Linv[i][c] = -(sum_{k=c}^{i-1} L[i][k] Linv[k][c]) / L[i][i]
            Linv[i][c] = -mp.fdot(L[i][c:i], LinvT_cols[c]) / L[i][i]
        for c in range(i + 1):
            LinvT_cols[c].append(Linv[i][c])
    a = [mp.fdot(Linv[i][:i + 1], y[:i + 1]) for i in range(n)]
    cols = [[Linv[r][c] for r in range(c, n)] for c in range(n)]   # column c of Linv, rows c..n-1
    alpha = [mp.fdot(cols[c], a[c:]) for c in range(n)]
    lml = -mp.fdot(a, a) / 2 - mp.mpf(n) / 2 * mp.log(2 * mp.pi) - mp.fsum(mp.log(L[i][i]) for i in range(n))
    grad = []
    if want_grad:
        gsum = [mp.mpf(0)] * (P + 1)
        for i in range(n):
            for j in range(i + 1):
                kinv = mp.fdot(cols[i], cols[j][i - j:])          # sum_{r>=i} Linv[r][i] Linv[r][j]
                w = alpha[i] * alpha[j] - kinv
                c = 1 if i == j else 2
                for p in range(P):
                    gsum[p] += c * w * dK[p][i][j]
                if i == j:
                    gsum[P] += w
        grad = [g / 2 for g in gsum]
    xs = [mp.mpf(float(v)) for v in Xs[:, 0]]
    mean, var = [], []
    for s in xs:
        ks = [kernel_eval(spec, MathMP, x[i] - s, x[i], s, one)[0] for i in range(n)]
        am = [mp.fdot(Linv[i][:i + 1], ks[:i + 1]) for i in range(n)]
        mean.append(mp.fdot(am, a))
        var.append(kernel_eval(spec, MathMP, mp.mpf(0), s, s, one)[0] - mp.fdot(am, am))
    f = lambda v: float(v)
    return f(lml), np.array([f(g) for g in grad]), np.array([f(m) for m in mean])[:, None], np.array([f(v) for v in var])[:, None]


def main():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    G = np.load(os.path.join(ROOT, "tests", "golden", "reference_call_pattern.npz"))
    out = {}
    cases = []
    Xd, Yd, Xsd = G["aapl_d_X"], G["aapl_d_Y"], G["aapl_d_Xs"]
    for name in list(KERNELS)[:8]:
        cases.append((f"aapl_d|{name}", name, Xd, Yd, Xsd[::4], True))
    X3, Y3 = make_c1(1000)
    X3, Y3 = X3[350:650].copy(), Y3[350:650].copy()            # 300 consecutive days at the C1 spacing
    Xs3 = np.concatenate([X3[::25], X3[-1:] + (X3[1] - X3[0]) * np.arange(1, 6)[:, None]])
    cases.append(("c1_300|SE+Periodic(SE)", "SE+Periodic(SE)", X3, Y3, Xs3, True))
    X1, Y1 = make_c1(1000)
    Xs1 = np.concatenate([X1[::50], X1[-1:] + (X1[1] - X1[0]) * np.arange(1, 11)[:, None]])
    cases.append(("c1_1000|SE+Periodic(SE)", "SE+Periodic(SE)", X1, Y1, Xs1, False))
    for key, name, X, Y, Xs, use_mp in cases:
        out[key + "|X"], out[key + "|Y"], out[key + "|Xs"] = X, Y, Xs
        for tag, s2 in (("1e-5", 1e-5), ("1e-2", 1e-2)):
            t0 = time.time()
            lml, g, m, v = gp_longdouble(KERNELS[name], X, Y, Xs, s2)
            t1 = time.time()
            k = f"{key}|{tag}"
            out[k + "|ld|lml"], out[k + "|ld|grad"], out[k + "|ld|mean"], out[k + "|ld|var"] = np.array(lml), g, m[:, None], v[:, None]
            msg = f"{k}: longdouble {t1 - t0:.1f}s lml={lml:.17g}"
            if use_mp:
                lm, gm, mm, vm = gp_mpmath(KERNELS[name], X, Y, Xs, s2)
                out[k + "|mp|lml"], out[k + "|mp|grad"], out[k + "|mp|mean"], out[k + "|mp|var"] = np.array(lm), gm, mm, vm
                msg += (f" | mpmath {time.time() - t1:.1f}s lml={lm:.17g} rel(ld-mp)={abs(lml - lm) / abs(lm):.1e} "
                        f"grad abs(ld-mp)={np.max(np.abs(g - gm)):.1e} var abs={np.max(np.abs(v[:, None] - vm)):.1e}")
            print(msg, flush=True)
    path = os.path.join(ROOT, "tests", "golden", "extended_precision_truth.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
