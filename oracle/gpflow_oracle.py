"""CPU oracle: NumPy/SciPy fp64 restatement of the GPflow 2.9.1 arithmetic that PortfolioOptGP drives.

TEST INFRASTRUCTURE ONLY.  Nothing under ``portfoliooptgp_b200/`` may import this module; it is
used by ``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, always as the checker or the timed CPU stand-in, never as a product path.

PARITY UNPINNED.  The arithmetic of the hot path lives in a third-party dependency that is not
vendored in the reference tree and cannot be installed here (no network):
``gpflow==2.9.1`` (reference ``Multi-Input_GPR/requirements.txt:37``) on ``tensorflow==2.16.1``
(``:108,111``), ``tensorflow-probability==0.24.0`` (``:112``).  The reference's own tests mock
GPflow (``GPR/tests/test_model_trainer.py:11-15``, ``GPR/tests/test_predictor.py:11-13``) and hold
no golden vectors.  This file therefore restates GPflow's *published* op sequence (module and
function named beside each routine below) and is pinned only by
closed-form known answers, extended-precision finite differences and the SVGP<->GPR identity
(``tests/test_oracle.py``), not by outputs of GPflow itself.  The exact-GP part (log marginal
likelihood, its gradient, predictive moments; every kernel family, sums, products, Periodic, Linear) is
additionally checked against an independent third-party implementation that IS installed here,
scikit-learn's GaussianProcessRegressor (``tests/test_oracle_sklearn.py``: 1e-10 on the LML, 2e-7 on
gradients, 1e-8 on predictions).

Reference call sites the routines serve (all paths relative to /root/reference):
  GPR/model_trainer.py:15-20            GPR(data, kernel); training_loss; predict_f
  GPR/predictor.py:6-7                  predict_f(full_cov=False); predict_y
  Multi-Input_GPR/models/model_trainer.py:19-21,31-40
  Multi-Input_GPR/main.py:126-135       k1(active_dims=slice) * k2(active_dims=slice)
  test_scripts/SVGP.py:515-540          SVGP(kernel, Gaussian(1e-4), Z, num_data); elbo; predict_f
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import scipy.linalg as sla

DEFAULT_JITTER = 1e-6  # gpflow.config.default_jitter()
VARIANCE_LOWER_BOUND = 1e-6  # gpflow.likelihoods.Gaussian DEFAULT_VARIANCE_LOWER_BOUND

# ----------------------------------------------------------------------------------------------
# G1  gpflow/base.py Parameter + gpflow/utilities/bijectors.py positive()
# ----------------------------------------------------------------------------------------------


def softplus(u):
    """tfp.bijectors.Softplus forward: log(1 + exp(u)), evaluated without overflow."""
    u = np.asarray(u, dtype=np.float64)
    return np.where(u > 0, u + np.log1p(np.exp(-np.abs(u))), np.log1p(np.exp(-np.abs(u))))


def softplus_inverse(theta):
    """Softplus inverse: log(expm1(theta)) = theta + log(1 - exp(-theta))."""
    theta = np.asarray(theta, dtype=np.float64)
    return theta + np.log(-np.expm1(-theta))


def sigmoid(u):
    """d softplus / du."""
    u = np.asarray(u, dtype=np.float64)
    return np.where(u >= 0, 1.0 / (1.0 + np.exp(-np.abs(u))), np.exp(-np.abs(u)) / (1.0 + np.exp(-np.abs(u))))


# ----------------------------------------------------------------------------------------------
# Kernel description (plain data; deliberately independent of the product package's classes)
# ----------------------------------------------------------------------------------------------

STATIONARY_KINDS = ("se", "rq", "matern12", "exponential", "matern32", "matern52")


@dataclass
class Leaf:
    """One GPflow leaf kernel.  kind in STATIONARY_KINDS + ("linear",).  Defaults = GPflow's."""

    kind: str
    variance: float = 1.0
    lengthscales: Union[float, Sequence[float]] = 1.0
    alpha: float = 1.0  # RationalQuadratic only
    active_dims: Optional[Union[slice, Sequence[int]]] = None


@dataclass
class Periodic:
    """gpflow.kernels.Periodic(base_kernel, period); active_dims are the base kernel's."""

    base: Leaf
    period: float = 1.0


@dataclass
class Sum:
    kernels: List = field(default_factory=list)


@dataclass
class Product:
    kernels: List = field(default_factory=list)


def _slice(X, active_dims):
    # gpflow/kernels/base.py Kernel.slice
    if active_dims is None:
        return X
    if isinstance(active_dims, slice):
        return X[..., active_dims]
    return X[..., np.asarray(active_dims, dtype=int)]


# G3  gpflow/utilities/ops.py square_distance (Gram form, may go slightly negative)
def square_distance(X, X2=None):
    if X2 is None:
        Xs = np.sum(np.square(X), axis=-1, keepdims=True)
        dist = -2.0 * (X @ X.T)
        dist += Xs + Xs.T
        return dist
    Xs = np.sum(np.square(X), axis=-1)
    X2s = np.sum(np.square(X2), axis=-1)
    dist = -2.0 * (X @ X2.T)
    dist += Xs[:, None] + X2s[None, :]
    return dist


def direct_square_distance(X, X2=None):
    """Direct-difference form sum_d (x_d - x'_d)^2 (what the fused CUDA kernel evaluates);
    mathematically identical to square_distance, used to quantify the Gram-form rounding gap."""
    X2 = X if X2 is None else X2
    out = np.zeros((X.shape[0], X2.shape[0]), dtype=np.float64)
    for j in range(X.shape[1]):      # one dimension at a time: O(N N2) memory at the BASELINE sizes
        d = X[:, j][:, None] - X2[:, j][None, :]
        d *= d
        out += d
    return out


_DISTANCE_FORM = "gram"  # "gram" = GPflow-faithful; tests flip to "direct" to bound the gap


def set_distance_form(form: str):
    global _DISTANCE_FORM
    assert form in ("gram", "direct")
    _DISTANCE_FORM = form


def _sqdist(X, X2):
    return square_distance(X, X2) if _DISTANCE_FORM == "gram" else direct_square_distance(X, X2)


# G4  gpflow/kernels/stationaries.py
def _k_r2(kind, variance, r2, alpha=1.0):
    if kind == "se":
        return variance * np.exp(-0.5 * r2)
    if kind == "rq":
        return variance * (1.0 + 0.5 * r2 / alpha) ** (-alpha)
    r = np.sqrt(np.maximum(r2, 1e-36))
    return _k_r(kind, variance, r)


def _k_r(kind, variance, r):
    if kind == "matern12":
        return variance * np.exp(-r)
    if kind == "exponential":
        return variance * np.exp(-0.5 * r)
    if kind == "matern32":
        s3 = np.sqrt(3.0)
        return variance * (1.0 + s3 * r) * np.exp(-s3 * r)
    if kind == "matern52":
        s5 = np.sqrt(5.0)
        return variance * (1.0 + s5 * r + 5.0 / 3.0 * np.square(r)) * np.exp(-s5 * r)
    raise ValueError(kind)


def K(kernel, X, X2=None):
    """kernel(X, X2) as GPflow's Kernel.__call__(full_cov=True): slice, then K.  G2-G6."""
    X = np.asarray(X, dtype=np.float64)
    X2 = None if X2 is None else np.asarray(X2, dtype=np.float64)
    if isinstance(kernel, Sum):  # gpflow/kernels/base.py Sum._reduce = tf.add_n
        out = K(kernel.kernels[0], X, X2)
        for k in kernel.kernels[1:]:
            out = out + K(k, X, X2)
        return out
    if isinstance(kernel, Product):  # Product._reduce = tf.reduce_prod over the stack
        out = K(kernel.kernels[0], X, X2)
        for k in kernel.kernels[1:]:
            out = out * K(k, X, X2)
        return out
    if isinstance(kernel, Periodic):  # G6 gpflow/kernels/periodic.py
        b = kernel.base
        Xs = _slice(X, b.active_dims)
        X2s = Xs if X2 is None else _slice(X2, b.active_dims)
        diff = Xs[:, None, :] - X2s[None, :, :]  # difference_matrix
        r = np.pi * diff / kernel.period
        scaled_sine = np.sin(r) / np.asarray(b.lengthscales, dtype=np.float64)
        if b.kind in ("se", "rq"):
            return _k_r2(b.kind, b.variance, np.sum(np.square(scaled_sine), -1), b.alpha)
        return _k_r(b.kind, b.variance, np.sum(np.abs(scaled_sine), -1))
    Xs = _slice(X, kernel.active_dims)
    X2s = None if X2 is None else _slice(X2, kernel.active_dims)
    if kernel.kind == "linear":  # G5 gpflow/kernels/linears.py
        return (Xs * kernel.variance) @ (Xs if X2s is None else X2s).T
    ls = np.asarray(kernel.lengthscales, dtype=np.float64)
    r2 = _sqdist(Xs / ls, None if X2s is None else X2s / ls)
    return _k_r2(kernel.kind, kernel.variance, r2, kernel.alpha)


def K_diag(kernel, X):
    X = np.asarray(X, dtype=np.float64)
    if isinstance(kernel, Sum):
        return sum(K_diag(k, X) for k in kernel.kernels)
    if isinstance(kernel, Product):
        out = K_diag(kernel.kernels[0], X)
        for k in kernel.kernels[1:]:
            out = out * K_diag(k, X)
        return out
    if isinstance(kernel, Periodic):
        return np.full(X.shape[0], float(kernel.base.variance))
    if kernel.kind == "linear":
        Xs = _slice(X, kernel.active_dims)
        return np.sum(np.square(Xs) * kernel.variance, axis=-1)
    return np.full(X.shape[0], float(kernel.variance))


# ----------------------------------------------------------------------------------------------
# Parameter enumeration in GPflow's trainable_variables order (tf.Module attribute traversal:
# attributes sorted by name at each level, list items by index) -- G10
# ----------------------------------------------------------------------------------------------


def kernel_params(kernel, prefix="kernel") -> List[Tuple[str, object, str]]:
    """[(path, owner, attribute)] in GPflow order; value = getattr(owner, attribute)."""
    if isinstance(kernel, (Sum, Product)):
        out = []
        for i, k in enumerate(kernel.kernels):
            out += kernel_params(k, f"{prefix}.kernels[{i}]")
        return out
    if isinstance(kernel, Periodic):
        return kernel_params(kernel.base, f"{prefix}.base_kernel") + [(f"{prefix}.period", kernel, "period")]
    out = []
    if kernel.kind == "rq":
        out.append((f"{prefix}.alpha", kernel, "alpha"))
    if kernel.kind != "linear":
        out.append((f"{prefix}.lengthscales", kernel, "lengthscales"))
    out.append((f"{prefix}.variance", kernel, "variance"))
    return out


def get_theta(kernel) -> np.ndarray:
    """Flat constrained parameter vector (ARD lengthscales expand in place)."""
    vals = []
    for _, owner, attr in kernel_params(kernel):
        vals.extend(np.atleast_1d(np.asarray(getattr(owner, attr), dtype=np.float64)).tolist())
    return np.asarray(vals, dtype=np.float64)


def set_theta(kernel, theta):
    theta = np.asarray(theta, dtype=np.float64)
    pos = 0
    for _, owner, attr in kernel_params(kernel):
        cur = getattr(owner, attr)
        n = np.size(cur)
        if np.ndim(cur) == 0:
            setattr(owner, attr, float(theta[pos]))
        else:
            setattr(owner, attr, theta[pos:pos + n].copy())
        pos += n
    assert pos == theta.size


# ----------------------------------------------------------------------------------------------
# G7-G9  gpflow/models/gpr.py GPR.log_marginal_likelihood, gpflow/logdensities.py
# ----------------------------------------------------------------------------------------------


def gpr_cholesky(kernel, X, noise_variance):
    Kmat = K(kernel, X)
    Kmat[np.diag_indices_from(Kmat)] += noise_variance  # add_noise_cov: set_diag(K, diag + s)
    return sla.cholesky(Kmat, lower=True, check_finite=False)


def multivariate_normal(d, L):
    """logdensities.multivariate_normal with x - mu = d [N,R]; returns [R]."""
    alpha = sla.solve_triangular(L, d, lower=True, check_finite=False)
    n = d.shape[0]
    p = -0.5 * np.sum(np.square(alpha), 0)
    p -= 0.5 * n * np.log(2 * np.pi)
    p -= np.sum(np.log(np.diag(L)))
    return p


def gpr_lml(kernel, X, Y, noise_variance, mean=None):
    Y = np.asarray(Y, dtype=np.float64).reshape(len(Y), -1)
    d = Y if mean is None else Y - mean
    L = gpr_cholesky(kernel, X, noise_variance)
    return float(np.sum(multivariate_normal(d, L)))


def gpr_training_loss(kernel, X, Y, noise_variance):
    """training_loss = -(LML + log prior); no priors on this path (G9)."""
    return -gpr_lml(kernel, X, Y, noise_variance)


def dK_dtheta(kernel, X) -> List[np.ndarray]:
    """Explicit dK/dtheta_p for every constrained parameter, in get_theta order (oracle only:
    the product never materialises these).  Central differences are NOT used: each entry is the
    closed-form derivative of the G4-G6 formulae."""
    X = np.asarray(X, dtype=np.float64)
    if isinstance(kernel, Sum):
        out = []
        for k in kernel.kernels:
            out += dK_dtheta(k, X)
        return out
    if isinstance(kernel, Product):
        Ks = [K(k, X) for k in kernel.kernels]
        out = []
        for i, k in enumerate(kernel.kernels):
            others = np.ones_like(Ks[0])
            for j, Kj in enumerate(Ks):
                if j != i:
                    others = others * Kj
            out += [g * others for g in dK_dtheta(k, X)]
        return out
    if isinstance(kernel, Periodic):
        b = kernel.base
        Xs = _slice(X, b.active_dims)
        diff = Xs[:, None, :] - Xs[None, :, :]
        arg = np.pi * diff / kernel.period
        ls = np.broadcast_to(np.asarray(b.lengthscales, dtype=np.float64), (Xs.shape[1],))
        sn, cs = np.sin(arg), np.cos(arg)
        darg_dp = -arg / kernel.period
        if b.kind in ("se", "rq"):
            s = np.sum(np.square(sn / ls), -1)
            dk_ds = _dk_dr2(b.kind, b.variance, s, b.alpha)
            per_dim = np.square(sn / ls)  # contribution of each dim to s
            ds_dp = np.sum(2.0 * sn * cs * darg_dp / np.square(ls), -1)
        else:
            s = np.sum(np.abs(sn / ls), -1)
            dk_ds = _dk_dr(b.kind, b.variance, s)
            per_dim = np.abs(sn / ls)
            ds_dp = np.sum(np.sign(sn) * cs * darg_dp / ls, -1)
        out = []
        if b.kind == "rq":
            out.append(_dk_dalpha_rq(b.variance, s, b.alpha))
        power = 2.0 if b.kind in ("se", "rq") else 1.0
        if np.ndim(b.lengthscales) == 0:
            out.append(dk_ds * (-power * s / float(b.lengthscales)))
        else:
            for d_ in range(Xs.shape[1]):
                out.append(dk_ds * (-power * per_dim[..., d_] / ls[d_]))
        kval = K(kernel, X)
        out.append(kval / b.variance)
        out.append(dk_ds * ds_dp)
        return out
    Xs = _slice(X, kernel.active_dims)
    if kernel.kind == "linear":
        if np.ndim(kernel.variance) == 0:
            return [Xs @ Xs.T]
        return [np.outer(Xs[:, d_], Xs[:, d_]) for d_ in range(Xs.shape[1])]
    ls = np.broadcast_to(np.asarray(kernel.lengthscales, dtype=np.float64), (Xs.shape[1],))
    if np.ndim(kernel.lengthscales) == 0:
        per_dim = None  # isotropic: only the [N,N] distance matrix is needed (no [N,N,D] tensor)
        r2 = np.maximum(_sqdist(Xs / ls, None), 0.0)
    else:
        diff = Xs[:, None, :] - Xs[None, :, :]
        per_dim = np.square(diff / ls)
        r2 = np.sum(per_dim, -1)
    dk = _dk_dr2(kernel.kind, kernel.variance, r2, kernel.alpha)
    out = []
    if kernel.kind == "rq":
        out.append(_dk_dalpha_rq(kernel.variance, r2, kernel.alpha))
    if np.ndim(kernel.lengthscales) == 0:
        out.append(dk * (-2.0 * r2 / float(kernel.lengthscales)))
    else:
        for d_ in range(Xs.shape[1]):
            out.append(dk * (-2.0 * per_dim[..., d_] / ls[d_]))
    out.append(_k_r2(kernel.kind, 1.0, r2, kernel.alpha))
    return out


def _dk_dr2(kind, variance, r2, alpha=1.0):
    """d k / d r2 (r2 = scaled squared distance), with the r -> 0 limits taken analytically."""
    if kind == "se":
        return -0.5 * variance * np.exp(-0.5 * r2)
    if kind == "rq":
        return -0.5 * variance * (1.0 + 0.5 * r2 / alpha) ** (-alpha - 1.0)
    r = np.sqrt(np.maximum(r2, 0.0))
    if kind == "matern32":
        return -1.5 * variance * np.exp(-np.sqrt(3.0) * r)
    if kind == "matern52":
        s5 = np.sqrt(5.0)
        return -(5.0 / 6.0) * variance * (1.0 + s5 * r) * np.exp(-s5 * r)
    # matern12 / exponential: dk/dr2 = dk/dr / (2r) is singular at r = 0, but dk/dlengthscale =
    # dk/dr2 * (-2 r2 / l) -> 0 there; return 0 so the product is 0 (GPflow's clamp gives the same).
    with np.errstate(divide="ignore", invalid="ignore"):
        c = 1.0 if kind == "matern12" else 0.5
        val = -c * variance * np.exp(-c * r) / (2.0 * r)
    return np.where(r > 0, val, 0.0)


def _dk_dr(kind, variance, r):
    if kind == "matern12":
        return -variance * np.exp(-r)
    if kind == "exponential":
        return -0.5 * variance * np.exp(-0.5 * r)
    if kind == "matern32":
        return -3.0 * variance * r * np.exp(-np.sqrt(3.0) * r)
    if kind == "matern52":
        s5 = np.sqrt(5.0)
        return -(5.0 / 3.0) * variance * r * (1.0 + s5 * r) * np.exp(-s5 * r)
    raise ValueError(kind)


def _dk_dalpha_rq(variance, r2, alpha):
    base = 1.0 + 0.5 * r2 / alpha
    return variance * base ** (-alpha) * (-np.log(base) + 0.5 * r2 / (alpha * base))


def gpr_lml_and_grad(kernel, X, Y, noise_variance):
    """LML and its gradient w.r.t. the CONSTRAINED parameters: (lml, dlml/dtheta [P], dlml/dnoise).

    dLML/dtheta_p = 1/2 tr((alpha alpha^T - K^-1) dK/dtheta_p), the identity the TF autodiff of
    G8 evaluates (CholeskyGrad + MatrixTriangularSolveGrad)."""
    Y = np.asarray(Y, dtype=np.float64).reshape(len(Y), -1)
    L = gpr_cholesky(kernel, X, noise_variance)
    lml = float(np.sum(multivariate_normal(Y, L)))
    n = L.shape[0]
    Linv = sla.solve_triangular(L, np.eye(n), lower=True, check_finite=False)
    Kinv = Linv.T @ Linv
    a = Kinv @ Y
    W = a @ a.T - Y.shape[1] * Kinv
    grads = np.array([0.5 * np.sum(W * dK) for dK in dK_dtheta(kernel, X)])
    gnoise = 0.5 * np.trace(W)
    return lml, grads, float(gnoise)


# ----------------------------------------------------------------------------------------------
# G11-G12  gpflow/posteriors.py GPRPosterior, gpflow/conditionals/util.py base_conditional_with_lm
# ----------------------------------------------------------------------------------------------


def gpr_predict_f(kernel, X, Y, noise_variance, Xnew, full_cov=False):
    Y = np.asarray(Y, dtype=np.float64).reshape(len(Y), -1)
    Lm = gpr_cholesky(kernel, X, noise_variance)
    Kmn = K(kernel, X, Xnew)
    A = sla.solve_triangular(Lm, Kmn, lower=True, check_finite=False)
    if full_cov:
        fvar = K(kernel, Xnew) - A.T @ A
    else:
        fvar = K_diag(kernel, Xnew) - np.sum(np.square(A), 0)
    A = sla.solve_triangular(Lm.T, A, lower=False, check_finite=False)
    fmean = A.T @ Y
    if full_cov:
        return fmean, fvar[None]
    return fmean, np.tile(fvar[:, None], (1, Y.shape[1]))


def gpr_predict_y(kernel, X, Y, noise_variance, Xnew):
    m, v = gpr_predict_f(kernel, X, Y, noise_variance, Xnew)
    return m, v + noise_variance


# ----------------------------------------------------------------------------------------------
# G13-G14  gpflow/models/svgp.py, conditionals/conditionals.py, kullback_leiblers.py
# ----------------------------------------------------------------------------------------------


def gauss_kl(q_mu, q_sqrt, Kp=None):
    """KL[N(q_mu, q_sqrt q_sqrt^T) || N(0, Kp)] (Kp None => whitened prior N(0, I)).
    q_mu [M,L]; q_sqrt [L,M,M] lower-triangular."""
    M, Lr = q_mu.shape
    Lq = np.tril(q_sqrt)
    if Kp is None:
        alpha = q_mu
        trace = np.sum(np.square(Lq))
        logdet_p = 0.0
    else:
        Lp = sla.cholesky(Kp, lower=True, check_finite=False)
        alpha = sla.solve_triangular(Lp, q_mu, lower=True, check_finite=False)
        trace = sum(np.sum(np.square(sla.solve_triangular(Lp, Lq[i], lower=True, check_finite=False)))
                    for i in range(Lr))
        logdet_p = Lr * np.sum(np.log(np.square(np.diag(Lp))))
    mahalanobis = np.sum(np.square(alpha))
    constant = -float(M * Lr)
    logdet_q = np.sum(np.log(np.square(np.diagonal(Lq, axis1=-2, axis2=-1))))
    return 0.5 * (mahalanobis + constant - logdet_q + trace + logdet_p)


def svgp_predict_f(kernel, Z, q_mu, q_sqrt, Xnew, whiten=True):
    """conditional(Xnew, InducingPoints(Z), kernel, q_mu, q_sqrt, full_cov=False, white=whiten)."""
    Z = np.asarray(Z, dtype=np.float64)
    Kmm = K(kernel, Z) + DEFAULT_JITTER * np.eye(Z.shape[0])
    Kmn = K(kernel, Z, Xnew)
    Knn = K_diag(kernel, Xnew)
    Lm = sla.cholesky(Kmm, lower=True, check_finite=False)
    A = sla.solve_triangular(Lm, Kmn, lower=True, check_finite=False)
    fvar = Knn - np.sum(np.square(A), 0)
    if not whiten:
        A = sla.solve_triangular(Lm.T, A, lower=False, check_finite=False)
    fmean = A.T @ q_mu
    Lr = q_mu.shape[1]
    fvar = np.tile(fvar[None, :], (Lr, 1))
    for i in range(Lr):
        LTA = np.tril(q_sqrt[i]).T @ A
        fvar[i] += np.sum(np.square(LTA), 0)
    return fmean, fvar.T


def svgp_elbo(kernel, Z, q_mu, q_sqrt, noise_variance, X, Y, num_data=None, whiten=True):
    Y = np.asarray(Y, dtype=np.float64).reshape(len(Y), -1)
    if whiten:
        kl = gauss_kl(q_mu, q_sqrt)
    else:
        Z = np.asarray(Z, dtype=np.float64)
        kl = gauss_kl(q_mu, q_sqrt, K(kernel, Z) + DEFAULT_JITTER * np.eye(Z.shape[0]))
    fmean, fvar = svgp_predict_f(kernel, Z, q_mu, q_sqrt, X, whiten=whiten)
    # likelihoods.Gaussian._variational_expectations
    var_exp = np.sum(-0.5 * np.log(2 * np.pi) - 0.5 * np.log(noise_variance)
                     - 0.5 * (np.square(Y - fmean) + fvar) / noise_variance, axis=-1)
    scale = 1.0 if num_data is None else float(num_data) / X.shape[0]
    return float(np.sum(var_exp) * scale - kl)


# ---- SGPR: gpflow/models/sgpr.py (2.9.1) SGPR_deprecated._common_calculation / elbo / predict_f --------
# reference call site: test_scripts/SVGP.py:393-399 (SGPR(data, SquaredExponential(), inducing_variable=Z)
# trained by Scipy in plot_model, then predict_y)
def _sgpr_common(kernel, Z, noise_variance, X):
    Z = np.asarray(Z, dtype=np.float64)
    M = Z.shape[0]
    Kdiag = K_diag(kernel, X)
    kuf = K(kernel, Z, X)
    kuu = K(kernel, Z) + DEFAULT_JITTER * np.eye(M)
    L = sla.cholesky(kuu, lower=True, check_finite=False)
    sigma_sq = float(noise_variance)
    sigma = np.sqrt(sigma_sq)
    A = sla.solve_triangular(L, kuf, lower=True, check_finite=False) / sigma
    AAT = A @ A.T
    B = AAT + np.eye(M)
    LB = sla.cholesky(B, lower=True, check_finite=False)
    return Kdiag, L, sigma_sq, sigma, A, AAT, LB


def sgpr_elbo(kernel, Z, noise_variance, X, Y, mean=None):
    """SGPR.elbo = const + logdet_term + quad_term (one output column)."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64).reshape(len(X), -1)
    N, outdim = Y.shape
    Kdiag, L, sigma_sq, sigma, A, AAT, LB = _sgpr_common(kernel, Z, noise_variance, X)
    # logdet_term
    half_logdet_b = np.sum(np.log(np.diag(LB)))
    log_sigma_sq = N * np.log(sigma_sq)
    logdet_k = -outdim * (half_logdet_b + 0.5 * log_sigma_sq)
    trace_k = np.sum(Kdiag / sigma_sq)
    trace_q = np.trace(AAT)
    logdet = logdet_k + 0.5 * outdim * (trace_q - trace_k)
    # quad_term
    err = Y - (0.0 if mean is None else np.asarray(mean, dtype=np.float64).reshape(N, -1))
    Aerr = A @ (err / sigma)
    c = sla.solve_triangular(LB, Aerr, lower=True, check_finite=False)
    quad = -0.5 * (np.sum(np.square(err) / sigma_sq) - np.sum(np.square(c)))
    const = -0.5 * N * outdim * np.log(2 * np.pi)
    return float(const + logdet + quad)


def sgpr_predict_f(kernel, Z, noise_variance, X, Y, Xnew, mean=None):
    """SGPR.predict_f(Xnew, full_cov=False) without the mean function added back."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64).reshape(len(X), -1)
    Kdiag, L, sigma_sq, sigma, A, AAT, LB = _sgpr_common(kernel, Z, noise_variance, X)
    err = Y - (0.0 if mean is None else np.asarray(mean, dtype=np.float64).reshape(len(X), -1))
    Kus = K(kernel, np.asarray(Z, dtype=np.float64), Xnew)
    Aerr = A @ err
    c = sla.solve_triangular(LB, Aerr, lower=True, check_finite=False) / sigma
    tmp1 = sla.solve_triangular(L, Kus, lower=True, check_finite=False)
    tmp2 = sla.solve_triangular(LB, tmp1, lower=True, check_finite=False)
    fmean = tmp2.T @ c
    fvar = K_diag(kernel, Xnew) + np.sum(np.square(tmp2), 0) - np.sum(np.square(tmp1), 0)
    return fmean, np.tile(fvar[:, None], (1, Y.shape[1]))
