"""GPU, BASELINE full sizes (C2: N = 8192, D = 8): properties that do not need the CPU oracle --
row-permutation invariance of the LML, the gradient against central differences of the LML
itself, predict_y - predict_f = sigma^2 -- and one independent cross-check of the LML against
cuSOLVER (torch.linalg.cholesky on the assembled matrix; a comparison point, never on the product path)."""
import math

import numpy as np
import pytest

from tests.helpers import make_multi_input

pytestmark = pytest.mark.gpu

N, D, NOISE = 8192, 8, 1e-2


def _kernel(gp):
    K = gp.kernels
    return K.SquaredExponential(lengthscales=1.3) + K.Matern52(variance=0.7, lengthscales=2.0) + K.Linear(variance=0.2)


def test_c2_lml_properties_and_fd_gradient(gp):
    import torch
    X, Y = make_multi_input(2, N, D)
    k = _kernel(gp)
    m = gp.models.GPR((X, Y), kernel=k, noise_variance=NOISE)
    lml, g, gn = m.lml_and_constrained_grads()
    # (1) permutation invariance
    perm = np.random.default_rng(0).permutation(N)
    m2 = gp.models.GPR((X[perm], Y[perm]), kernel=k, noise_variance=NOISE)
    assert abs(float(m2.log_marginal_likelihood()) - lml) <= 1e-9 * abs(lml)
    # (2) independent LML through cuSOLVER on the engine-assembled K
    from portfoliooptgp_b200 import ops
    Kfull = ops.kernel_matrix(k, X, diag_add=NOISE)
    L = torch.linalg.cholesky(Kfull)
    a = torch.linalg.solve_triangular(L, torch.as_tensor(Y, device="cuda"), upper=False)
    ref = float(-0.5 * (a * a).sum() - 0.5 * N * math.log(2 * math.pi) - torch.log(torch.diagonal(L)).sum())
    assert abs(lml - ref) <= 1e-9 * abs(ref)
    # (3) directional derivative of the LML itself (constrained parameters + noise)
    theta0 = m._compiled.theta().copy() if m._compiled is not None else None
    params = list(k.parameters)
    direction = np.random.default_rng(1).standard_normal(len(params) + 1)
    h = 1e-5

    def lml_at(t):
        for p, v0, d in zip(params, base, direction[:-1]):
            p.assign(v0 * (1.0 + t * d * 0.1))
        m.likelihood.variance.assign(NOISE * (1.0 + t * direction[-1] * 0.1))
        return float(m.log_marginal_likelihood())

    base = [float(p.numpy()) for p in params]
    fd = (lml_at(h) - lml_at(-h)) / (2 * h)
    lml_at(0.0)
    want = sum(gi * v0 * d * 0.1 for gi, v0, d in zip(g, base, direction[:-1])) + gn * NOISE * direction[-1] * 0.1
    assert abs(fd - want) <= 1e-5 * max(1.0, abs(want))


def test_c2_predict_properties(gp):
    X, Y = make_multi_input(2, N, D)
    Xs, _ = make_multi_input(3, 1000, D)
    m = gp.models.GPR((X, Y), kernel=_kernel(gp), noise_variance=NOISE)
    fm, fv = m.predict_f(Xs)
    ym, yv = m.predict_y(Xs)
    assert np.array_equal(fm.numpy(), ym.numpy())
    assert np.allclose(yv.numpy() - fv.numpy(), NOISE, rtol=0, atol=1e-14)
    assert np.all(fv.numpy() > 0) and np.all(np.isfinite(fm.numpy()))
    # posterior variance at training inputs is below the noise variance
    _, tv = m.predict_f(X[:512])
    assert np.all(tv.numpy() < NOISE) and np.all(tv.numpy() > 0)
