"""Parameter / Module: host-side mirror of gpflow/base.py.

* ``Parameter`` holds an UNCONSTRAINED fp64 array (exposed as ``unconstrained_variable``, what
  ``model.trainable_variables`` yields and what ``optimizers.Scipy`` packs) plus a bijector;
  the constrained value (what kernels read) is ``transform.forward(u)``.
* positive parameters use softplus (tfp.bijectors.Softplus), optionally shifted by a lower bound
  (``gpflow.utilities.positive(lower=...)``; ``Gaussian.variance`` uses 1e-6).
* ``Module`` reproduces tf.Module's attribute traversal order (attributes sorted by name at each
  level, list items by index), which fixes the order of ``trainable_variables`` and therefore the
  layout of the L-BFGS vector (G10).
* Objects have reference semantics: the same kernel instance may sit in several models
  (reference GPR/model_trainer.py:15 reuses the instances created at GPR/main.py:105-114), so
  values are read from the Parameter objects at every evaluation; ``copy.deepcopy`` works
  (Multi-Input_GPR/main.py:176, models/model_trainer.py:31).
"""
from __future__ import annotations

import math
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np

DEFAULT_FLOAT = np.float64


# ---- bijectors -------------------------------------------------------------------------------


class Identity:
    name = "Identity"

    def forward(self, u):
        return u

    def inverse(self, x):
        return x

    def forward_grad(self, u):
        """d forward / du, elementwise."""
        return np.ones_like(u)

    def __repr__(self):
        return "Identity"


class Softplus:
    """theta = log(1 + exp(u)) + lower   (tfp Softplus chained with Shift(lower))."""

    def __init__(self, lower: float = 0.0):
        self.lower = float(lower)

    @property
    def name(self):
        return "Softplus" if self.lower == 0.0 else "Softplus + Shift"

    def forward(self, u):
        u = np.asarray(u, dtype=DEFAULT_FLOAT)
        return np.maximum(u, 0.0) + np.log1p(np.exp(-np.abs(u))) + self.lower

    def inverse(self, x):
        x = np.asarray(x, dtype=DEFAULT_FLOAT) - self.lower
        if np.any(x <= 0.0):
            raise ValueError(f"value must be > {self.lower} for a positive parameter, got {x + self.lower}")
        return x + np.log(-np.expm1(-x))

    def forward_grad(self, u):
        u = np.asarray(u, dtype=DEFAULT_FLOAT)
        e = np.exp(-np.abs(u))
        return np.where(u >= 0, 1.0 / (1.0 + e), e / (1.0 + e))

    def __repr__(self):
        return self.name


class FillTriangular:
    """gpflow.utilities.triangular(): unconstrained vector of length M(M+1)/2 per latent <-> lower
    triangular [L, M, M].  (tfp FillTriangular uses a spiral packing order; the order only permutes
    the L-BFGS vector, which leaves the iterates unchanged up to rounding, so row-major packing of
    the lower triangle is used here -- SURVEY.md H8.)"""

    name = "FillTriangular"

    def __init__(self, M: int):
        self.M = int(M)
        self._idx = np.tril_indices(self.M)

    def forward(self, u):
        u = np.asarray(u, dtype=DEFAULT_FLOAT)
        Lr = u.shape[0]
        out = np.zeros((Lr, self.M, self.M), dtype=DEFAULT_FLOAT)
        out[:, self._idx[0], self._idx[1]] = u
        return out

    def inverse(self, x):
        x = np.asarray(x, dtype=DEFAULT_FLOAT)
        return x[:, self._idx[0], self._idx[1]].copy()

    def forward_grad(self, u):
        return np.ones_like(u)

    def pull_back(self, g_full):
        """gradient w.r.t. the full [L,M,M] array -> gradient w.r.t. the packed vector."""
        return np.asarray(g_full)[:, self._idx[0], self._idx[1]].copy()

    def __repr__(self):
        return self.name


def positive(lower: Optional[float] = None) -> Softplus:
    """gpflow.utilities.positive: softplus, shifted by ``lower`` when given."""
    return Softplus(0.0 if lower is None else lower)


def triangular(M: int) -> FillTriangular:
    return FillTriangular(M)


# ---- variables / parameters ----------------------------------------------------------------------


class Variable:
    """Stand-in for the tf.Variable GPflow hands to the optimiser: an unconstrained fp64 array."""

    def __init__(self, value, name: str = "Variable"):
        self._value = np.array(value, dtype=DEFAULT_FLOAT)
        self.name = name
        self.version = 0

    def numpy(self) -> np.ndarray:
        return self._value.copy()

    def assign(self, value):
        value = np.asarray(value, dtype=DEFAULT_FLOAT)
        if value.shape != self._value.shape:
            value = value.reshape(self._value.shape)
        self._value = value.copy()
        self.version += 1
        return self

    @property
    def shape(self) -> Tuple[int, ...]:
        return self._value.shape

    @property
    def size(self) -> int:
        return int(self._value.size)

    @property
    def dtype(self):
        return self._value.dtype

    def __array__(self, dtype=None, copy=None):
        return self._value.astype(dtype) if dtype is not None else self._value.copy()

    def __repr__(self):
        return f"<Variable {self.name} shape={self.shape} value={self._value}>"


class Parameter:
    """gpflow.Parameter: constrained view over an unconstrained Variable."""

    def __init__(self, value, transform=None, trainable: bool = True, name: Optional[str] = None, prior=None,
                 dtype=None):
        if isinstance(value, Parameter):
            transform = transform if transform is not None else value.transform
            trainable = value.trainable if trainable is None else trainable
            value = value.numpy()
        if prior is not None:
            raise NotImplementedError("priors are not used on the PortfolioOptGP path (SURVEY.md G9)")
        self.transform = transform if transform is not None else Identity()
        value = np.asarray(_to_numpy(value), dtype=DEFAULT_FLOAT)
        self.unconstrained_variable = Variable(self.transform.inverse(value), name=name or "Parameter")
        self.trainable = bool(trainable)
        self.prior = None
        self.name = name

    # value access ------------------------------------------------------------------------------
    def numpy(self) -> np.ndarray:
        return np.asarray(self.transform.forward(self.unconstrained_variable._value), dtype=DEFAULT_FLOAT)

    def value(self):
        return self.numpy()

    def assign(self, value):
        """Assign a CONSTRAINED value (reference GPR/model_trainer.py:16)."""
        value = np.asarray(_to_numpy(value), dtype=DEFAULT_FLOAT)
        if value.shape != self.shape:
            value = np.broadcast_to(value, self.shape)
        u = self.transform.inverse(value)
        self.unconstrained_variable.assign(u)
        return self

    @property
    def shape(self):
        return self.unconstrained_variable.shape if not isinstance(self.transform, FillTriangular) else \
            (self.unconstrained_variable.shape[0], self.transform.M, self.transform.M)

    @property
    def dtype(self):
        return DEFAULT_FLOAT

    @property
    def size(self) -> int:
        return int(np.prod(self.shape)) if self.shape else 1

    def __array__(self, dtype=None, copy=None):
        v = self.numpy()
        return v.astype(dtype) if dtype is not None else v

    def __float__(self):
        return float(self.numpy())

    def __repr__(self):
        return f"<Parameter {self.name or ''} transform={self.transform!r} trainable={self.trainable} value={self.numpy()}>"

    # arithmetic on the constrained value (reference code multiplies/prints parameters)
    def _binop(self, other, op):
        return op(self.numpy(), _to_numpy(other))

    def __add__(self, o): return self._binop(o, np.add)
    def __radd__(self, o): return self._binop(o, lambda a, b: np.add(b, a))
    def __sub__(self, o): return self._binop(o, np.subtract)
    def __rsub__(self, o): return self._binop(o, lambda a, b: np.subtract(b, a))
    def __mul__(self, o): return self._binop(o, np.multiply)
    def __rmul__(self, o): return self._binop(o, lambda a, b: np.multiply(b, a))
    def __truediv__(self, o): return self._binop(o, np.divide)
    def __rtruediv__(self, o): return self._binop(o, lambda a, b: np.divide(b, a))
    def __pow__(self, o): return self._binop(o, np.power)
    def __neg__(self): return -self.numpy()
    def __getitem__(self, i): return self.numpy()[i]


def _to_numpy(x):
    if isinstance(x, (Parameter, Variable)):
        return x.numpy()
    if hasattr(x, "detach") and hasattr(x, "cpu"):  # torch tensor
        return x.detach().cpu().numpy()
    if hasattr(x, "numpy") and not isinstance(x, np.ndarray):
        return x.numpy()
    return x


# ---- module --------------------------------------------------------------------------------------


class Module:
    """tf.Module-like container: discovers Parameters / sub-Modules through its attributes."""

    def _children(self) -> Iterator[Tuple[str, object]]:
        for key in sorted(vars(self).keys()):
            if key.startswith("_"):
                continue
            yield key, vars(self)[key]

    def _flatten(self, prefix: str, seen: set) -> Iterator[Tuple[str, Parameter]]:
        for key, val in self._children():
            yield from _flatten_value(f"{prefix}.{key}" if prefix else key, val, seen)

    def named_parameters(self) -> List[Tuple[str, Parameter]]:
        return list(self._flatten("", set()))

    @property
    def parameters(self) -> Tuple[Parameter, ...]:
        return tuple(p for _, p in self.named_parameters())

    @property
    def trainable_parameters(self) -> Tuple[Parameter, ...]:
        return tuple(p for p in self.parameters if p.trainable)

    @property
    def variables(self) -> Tuple[Variable, ...]:
        return tuple(p.unconstrained_variable for p in self.parameters)

    @property
    def trainable_variables(self) -> Tuple[Variable, ...]:
        return tuple(p.unconstrained_variable for p in self.parameters if p.trainable)

    @property
    def submodules(self) -> Tuple["Module", ...]:
        out: List[Module] = []
        seen = set()

        def walk(v):
            if isinstance(v, Module):
                if id(v) in seen:
                    return
                seen.add(id(v))
                out.append(v)
                for _, c in v._children():
                    walk(c)
            elif isinstance(v, (list, tuple)):
                for c in v:
                    walk(c)
            elif isinstance(v, dict):
                for k in sorted(v):
                    walk(v[k])

        for _, c in self._children():
            walk(c)
        return tuple(out)


def _flatten_value(path: str, val, seen: set) -> Iterator[Tuple[str, Parameter]]:
    if isinstance(val, Parameter):
        if id(val) not in seen:
            seen.add(id(val))
            yield path, val
    elif isinstance(val, Module):
        yield from val._flatten(path, seen)
    elif isinstance(val, (list, tuple)):
        for i, v in enumerate(val):
            yield from _flatten_value(f"{path}[{i}]", v, seen)
    elif isinstance(val, dict):
        for k in sorted(val):
            yield from _flatten_value(f"{path}['{k}']", val[k], seen)


def set_trainable(model, flag: bool) -> None:
    """gpflow.set_trainable: a Parameter, a Module (all of its parameters), or an iterable of
    them.  Reference: GPR/model_trainer.py:17, Multi-Input_GPR/models/model_trainer.py:19,34."""
    if isinstance(model, Parameter):
        model.trainable = bool(flag)
    elif isinstance(model, Module):
        for p in model.parameters:
            p.trainable = bool(flag)
    elif isinstance(model, (list, tuple)):
        for m in model:
            set_trainable(m, flag)
    else:
        raise TypeError(f"set_trainable: unsupported object {type(model)}")
