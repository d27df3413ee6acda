"""CPU oracle #2: the same GPflow 2.9.1 graph as ``gpflow_oracle.py`` written with torch CPU fp64
ops so that reverse-mode autodiff (``torch.autograd``) plays the role TensorFlow's GradientTape
plays inside ``gpflow.optimizers.Scipy``: gradients here come from
differentiating *through* cholesky / triangular_solve, not from the analytic trace identity, which
makes the two oracles independent checks of each other.

TEST INFRASTRUCTURE ONLY -- never imported by ``portfoliooptgp_b200``.  PARITY UNPINNED (see the
header of ``gpflow_oracle.py``: gpflow==2.9.1 is un-vendored and not installable here).
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from . import gpflow_oracle as O

_D = torch.float64


def _slice(X, active_dims):
    if active_dims is None:
        return X
    if isinstance(active_dims, slice):
        return X[..., active_dims]
    return X[..., torch.as_tensor(list(active_dims), dtype=torch.long)]


def square_distance(X, X2=None):  # gpflow/utilities/ops.py (G3)
    if X2 is None:
        Xs = torch.sum(torch.square(X), dim=-1, keepdim=True)
        return -2.0 * (X @ X.T) + Xs + Xs.T
    Xs = torch.sum(torch.square(X), dim=-1)
    X2s = torch.sum(torch.square(X2), dim=-1)
    return -2.0 * (X @ X2.T) + Xs[:, None] + X2s[None, :]


def direct_square_distance(X, X2=None):
    """sum_d (x_d - x'_d)^2 evaluated directly (see gpflow_oracle.direct_square_distance)."""
    X2 = X if X2 is None else X2
    d = X[:, None, :] - X2[None, :, :]
    return torch.sum(d * d, dim=-1)


def _sqdist(X, X2):
    # follows gpflow_oracle.set_distance_form so both oracles flip together
    return square_distance(X, X2) if O._DISTANCE_FORM == "gram" else direct_square_distance(X, X2)


def _k_r2(kind, variance, r2, alpha):
    if kind == "se":
        return variance * torch.exp(-0.5 * r2)
    if kind == "rq":
        return variance * (1.0 + 0.5 * r2 / alpha) ** (-alpha)
    r = torch.sqrt(torch.clamp(r2, min=1e-36))
    return _k_r(kind, variance, r)


def _k_r(kind, variance, r):
    if kind == "matern12":
        return variance * torch.exp(-r)
    if kind == "exponential":
        return variance * torch.exp(-0.5 * r)
    if kind == "matern32":
        s3 = math.sqrt(3.0)
        return variance * (1.0 + s3 * r) * torch.exp(-s3 * r)
    if kind == "matern52":
        s5 = math.sqrt(5.0)
        return variance * (1.0 + s5 * r + 5.0 / 3.0 * torch.square(r)) * torch.exp(-s5 * r)
    raise ValueError(kind)


class _Theta:
    """Walks a flat constrained-parameter tensor in O.kernel_params order."""

    def __init__(self, kernel, theta):
        self.map = {}
        pos = 0
        for _, owner, attr in O.kernel_params(kernel):
            n = int(np.size(getattr(owner, attr)))
            self.map[(id(owner), attr)] = theta[pos] if np.ndim(getattr(owner, attr)) == 0 else theta[pos:pos + n]
            pos += n
        assert pos == theta.numel()

    def get(self, owner, attr):
        return self.map[(id(owner), attr)]


def K(kernel, th: _Theta, X, X2=None):
    if isinstance(kernel, O.Sum):
        out = K(kernel.kernels[0], th, X, X2)
        for k in kernel.kernels[1:]:
            out = out + K(k, th, X, X2)
        return out
    if isinstance(kernel, O.Product):
        out = K(kernel.kernels[0], th, X, X2)
        for k in kernel.kernels[1:]:
            out = out * K(k, th, X, X2)
        return out
    if isinstance(kernel, O.Periodic):
        b = kernel.base
        Xs = _slice(X, b.active_dims)
        X2s = Xs if X2 is None else _slice(X2, b.active_dims)
        r = math.pi * (Xs[:, None, :] - X2s[None, :, :]) / th.get(kernel, "period")
        scaled_sine = torch.sin(r) / th.get(b, "lengthscales")
        alpha = th.get(b, "alpha") if b.kind == "rq" else None
        if b.kind in ("se", "rq"):
            return _k_r2(b.kind, th.get(b, "variance"), torch.sum(torch.square(scaled_sine), -1), alpha)
        return _k_r(b.kind, th.get(b, "variance"), torch.sum(torch.abs(scaled_sine), -1))
    Xs = _slice(X, kernel.active_dims)
    X2s = None if X2 is None else _slice(X2, kernel.active_dims)
    if kernel.kind == "linear":
        return (Xs * th.get(kernel, "variance")) @ (Xs if X2s is None else X2s).T
    ls = th.get(kernel, "lengthscales")
    r2 = _sqdist(Xs / ls, None if X2s is None else X2s / ls)
    alpha = th.get(kernel, "alpha") if kernel.kind == "rq" else None
    return _k_r2(kernel.kind, th.get(kernel, "variance"), r2, alpha)


def K_diag(kernel, th: _Theta, X):
    if isinstance(kernel, O.Sum):
        out = K_diag(kernel.kernels[0], th, X)
        for k in kernel.kernels[1:]:
            out = out + K_diag(k, th, X)
        return out
    if isinstance(kernel, O.Product):
        out = K_diag(kernel.kernels[0], th, X)
        for k in kernel.kernels[1:]:
            out = out * K_diag(k, th, X)
        return out
    if isinstance(kernel, O.Periodic):
        return th.get(kernel.base, "variance") * torch.ones(X.shape[0], dtype=_D)
    if kernel.kind == "linear":
        return torch.sum(torch.square(_slice(X, kernel.active_dims)) * th.get(kernel, "variance"), dim=-1)
    return th.get(kernel, "variance") * torch.ones(X.shape[0], dtype=_D)


def gpr_lml(kernel, theta, noise_variance, X, Y):
    """GPR.log_marginal_likelihood (G7-G8) as a differentiable torch scalar."""
    th = _Theta(kernel, theta)
    Kmat = K(kernel, th, X)
    ks = Kmat + noise_variance * torch.eye(X.shape[0], dtype=_D)
    L = torch.linalg.cholesky(ks)
    alpha = torch.linalg.solve_triangular(L, Y, upper=False)
    n = Y.shape[0]
    p = -0.5 * torch.sum(torch.square(alpha), 0)
    p = p - 0.5 * n * math.log(2 * math.pi)
    p = p - torch.sum(torch.log(torch.diagonal(L)))
    return torch.sum(p)


def gpr_lml_and_grad(kernel, X, Y, noise_variance):
    """(lml, dlml/dtheta (constrained, O.get_theta order), dlml/dnoise) by autograd."""
    theta = torch.tensor(O.get_theta(kernel), dtype=_D, requires_grad=True)
    nv = torch.tensor(float(noise_variance), dtype=_D, requires_grad=True)
    Xt = torch.as_tensor(np.asarray(X, dtype=np.float64))
    Yt = torch.as_tensor(np.asarray(Y, dtype=np.float64).reshape(len(Y), -1))
    lml = gpr_lml(kernel, theta, nv, Xt, Yt)
    g_theta, g_nv = torch.autograd.grad(lml, [theta, nv])
    return float(lml.detach()), g_theta.numpy().copy(), float(g_nv)


def gauss_kl_white(q_mu, q_sqrt):
    Lq = torch.tril(q_sqrt)
    M, Lr = q_mu.shape
    mahalanobis = torch.sum(torch.square(q_mu))
    logdet_q = torch.sum(torch.log(torch.square(torch.diagonal(Lq, dim1=-2, dim2=-1))))
    trace = torch.sum(torch.square(Lq))
    return 0.5 * (mahalanobis - float(M * Lr) - logdet_q + trace)


def svgp_elbo(kernel, theta, Z, q_mu, q_sqrt, noise_variance, X, Y, num_data: Optional[float] = None):
    """SVGP.elbo, whiten=True, q_diag=False (G13-G14), differentiable in theta, Z, q_mu, q_sqrt."""
    th = _Theta(kernel, theta)
    M = Z.shape[0]
    Kmm = K(kernel, th, Z) + O.DEFAULT_JITTER * torch.eye(M, dtype=_D)
    Kmn = K(kernel, th, Z, X)
    Knn = K_diag(kernel, th, X)
    Lm = torch.linalg.cholesky(Kmm)
    A = torch.linalg.solve_triangular(Lm, Kmn, upper=False)
    fvar = Knn - torch.sum(torch.square(A), 0)
    fmean = A.T @ q_mu
    fvar = fvar[None, :].repeat(q_mu.shape[1], 1)
    LTA = torch.tril(q_sqrt).transpose(-1, -2) @ A[None]
    fvar = (fvar + torch.sum(torch.square(LTA), -2)).T
    var_exp = torch.sum(-0.5 * math.log(2 * math.pi) - 0.5 * torch.log(noise_variance)
                        - 0.5 * (torch.square(Y - fmean) + fvar) / noise_variance, dim=-1)
    scale = 1.0 if num_data is None else float(num_data) / X.shape[0]
    return torch.sum(var_exp) * scale - gauss_kl_white(q_mu, q_sqrt)


def svgp_elbo_and_grad(kernel, Z, q_mu, q_sqrt, noise_variance, X, Y, num_data=None):
    """(elbo, dict of gradients w.r.t. constrained theta, Z, q_mu, tril(q_sqrt), noise)."""
    theta = torch.tensor(O.get_theta(kernel), dtype=_D, requires_grad=True)
    Zt = torch.tensor(np.asarray(Z, dtype=np.float64), requires_grad=True)
    qm = torch.tensor(np.asarray(q_mu, dtype=np.float64), requires_grad=True)
    qs = torch.tensor(np.asarray(q_sqrt, dtype=np.float64), requires_grad=True)
    nv = torch.tensor(float(noise_variance), dtype=_D, requires_grad=True)
    Xt = torch.as_tensor(np.asarray(X, dtype=np.float64))
    Yt = torch.as_tensor(np.asarray(Y, dtype=np.float64).reshape(len(Y), -1))
    elbo = svgp_elbo(kernel, theta, Zt, qm, qs, nv, Xt, Yt, num_data)
    g = torch.autograd.grad(elbo, [theta, Zt, qm, qs, nv])
    return float(elbo.detach()), {"theta": g[0].numpy().copy(), "Z": g[1].numpy().copy(), "q_mu": g[2].numpy().copy(),
                         "q_sqrt": np.tril(g[3].numpy()).copy(), "noise": float(g[4])}


def sgpr_elbo(kernel, theta, Z, noise_variance, X, err):
    """SGPR.elbo (gpflow/models/sgpr.py), differentiable in theta, Z, noise_variance and err = Y - m(X)."""
    th = _Theta(kernel, theta)
    M, N = Z.shape[0], X.shape[0]
    Kdiag = K_diag(kernel, th, X)
    kuf = K(kernel, th, Z, X)
    kuu = K(kernel, th, Z) + O.DEFAULT_JITTER * torch.eye(M, dtype=_D)
    L = torch.linalg.cholesky(kuu)
    sigma = torch.sqrt(noise_variance)
    A = torch.linalg.solve_triangular(L, kuf, upper=False) / sigma
    AAT = A @ A.T
    LB = torch.linalg.cholesky(AAT + torch.eye(M, dtype=_D))
    half_logdet_b = torch.sum(torch.log(torch.diagonal(LB)))
    logdet = -(half_logdet_b + 0.5 * N * torch.log(noise_variance)) + 0.5 * (torch.trace(AAT) - torch.sum(Kdiag / noise_variance))
    Aerr = A @ (err / sigma)
    c = torch.linalg.solve_triangular(LB, Aerr, upper=False)
    quad = -0.5 * (torch.sum(torch.square(err) / noise_variance) - torch.sum(torch.square(c)))
    return -0.5 * N * math.log(2 * math.pi) + logdet + quad


def sgpr_elbo_and_grad(kernel, Z, noise_variance, X, Y, mean=None):
    """(elbo, gradients w.r.t. constrained theta, Z, noise variance and err) by reverse-mode autodiff."""
    theta = torch.tensor(O.get_theta(kernel), dtype=_D, requires_grad=True)
    Zt = torch.tensor(np.asarray(Z, dtype=np.float64), requires_grad=True)
    nv = torch.tensor(float(noise_variance), dtype=_D, requires_grad=True)
    Xt = torch.as_tensor(np.asarray(X, dtype=np.float64))
    e = np.asarray(Y, dtype=np.float64).reshape(len(Y), 1)
    if mean is not None:
        e = e - np.asarray(mean, dtype=np.float64).reshape(len(Y), 1)
    et = torch.tensor(e, requires_grad=True)
    elbo = sgpr_elbo(kernel, theta, Zt, nv, Xt, et)
    g = torch.autograd.grad(elbo, [theta, Zt, nv, et])
    return float(elbo.detach()), {"theta": g[0].numpy().copy(), "Z": g[1].numpy().copy(), "noise": float(g[2]),
                                  "err": g[3].numpy()[:, 0].copy()}
