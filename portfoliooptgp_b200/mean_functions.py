"""Mean functions.  The hot-path call sites use the default (Zero); Constant / Linear appear in
the exploratory scripts (test_scripts/GPFlow.py:186-190, GPR_Class.py:101) and are evaluated on
the device with torch ops before the residual reaches the engine."""
from __future__ import annotations

import numpy as np
import torch

from .base import Module, Parameter


class MeanFunction(Module):
    def __call__(self, X: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    # d(sum_i g_i m(x_i)) / d parameters, as {id(param): unconstrained gradient}
    def backward(self, X: torch.Tensor, g: torch.Tensor):
        return {}


class Zero(MeanFunction):
    def __call__(self, X):
        return torch.zeros((X.shape[0], 1), dtype=X.dtype, device=X.device)


class Constant(MeanFunction):
    def __init__(self, c=None):
        self.c = Parameter(np.zeros(1) if c is None else np.atleast_1d(np.asarray(c, dtype=np.float64)), name="c")

    def __call__(self, X):
        c = torch.as_tensor(self.c.numpy(), dtype=X.dtype, device=X.device)
        return torch.ones((X.shape[0], 1), dtype=X.dtype, device=X.device) * c.reshape(1, -1)

    def backward(self, X, g):
        return {id(self.c): np.atleast_1d(float(g.sum().item()))}


class Linear(MeanFunction):
    def __init__(self, A=None, b=None):
        self.A = Parameter(np.ones((1, 1)) if A is None else np.asarray(A, dtype=np.float64), name="A")
        self.b = Parameter(np.zeros(1) if b is None else np.atleast_1d(np.asarray(b, dtype=np.float64)), name="b")

    def __call__(self, X):
        A = torch.as_tensor(self.A.numpy(), dtype=X.dtype, device=X.device)
        b = torch.as_tensor(self.b.numpy(), dtype=X.dtype, device=X.device)
        return X @ A + b

    def backward(self, X, g):
        gA = (X.T @ g.reshape(-1, 1)).cpu().numpy()
        return {id(self.A): gA.reshape(self.A.shape), id(self.b): np.atleast_1d(float(g.sum().item()))}


class Polynomial(MeanFunction):
    """gpflow.functions.Polynomial(degree): m(x) = sum_p w_p x^p over all cross terms of the input
    columns (test_scripts/GPR.py:103 uses Polynomial(2) on a single time column)."""

    def __init__(self, degree: int, w=None):
        self.degree = int(degree)
        self._w_init = w
        self.w = None  # created on first call, when the input dimension is known (GPflow does the same lazily)

    def _children(self):
        for key, val in super()._children():
            if key != "degree" and val is not None:
                yield key, val

    def _powers(self, D):
        import itertools
        return [p for p in itertools.product(range(self.degree + 1), repeat=D) if sum(p) <= self.degree]

    def _features(self, X):
        pw = self._powers(X.shape[1])
        cols = []
        for p in pw:
            c = torch.ones(X.shape[0], dtype=X.dtype, device=X.device)
            for d, e in enumerate(p):
                if e:
                    c = c * X[:, d] ** e
            cols.append(c)
        return torch.stack(cols, dim=1)

    def _ensure(self, D):
        if self.w is None:
            n = len(self._powers(D))
            w0 = np.zeros((n, 1)) if self._w_init is None else np.asarray(self._w_init, dtype=np.float64).reshape(n, 1)
            self.w = Parameter(w0, name="w")

    def __call__(self, X):
        self._ensure(X.shape[1])
        w = torch.as_tensor(self.w.numpy(), dtype=X.dtype, device=X.device)
        return self._features(X) @ w

    def backward(self, X, g):
        self._ensure(X.shape[1])
        return {id(self.w): (self._features(X).T @ g.reshape(-1, 1)).cpu().numpy()}
