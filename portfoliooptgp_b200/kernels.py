"""Covariance functions: host-side mirror of gpflow.kernels for the subset PortfolioOptGP uses
(reference GPR/main.py:105-114, Multi-Input_GPR/main.py:126-135,520-528, test_scripts/SVGP.py:496-506).

The objects only hold Parameters and structure.  Evaluation happens on the GPU: ``compile_kernel``
lowers an expression tree to the sum-of-products descriptor of include/gpb200.h
(``gpb_kernel_spec``), which the fused assembly / gradient kernels interpret per matrix element.

GPflow facts mirrored: defaults ``variance = lengthscales = period = alpha =
1.0``; ``a + b`` -> Sum (nested sums flattened), ``a * b`` -> Product; ``active_dims`` (slice or
index list) selects columns per leaf; Periodic takes its active_dims from the base kernel; the
attribute names (hence the trainable-variable order) are ``alpha, lengthscales, variance`` for
stationary kernels, ``variance`` for Linear, ``base_kernel, period`` for Periodic, ``kernels`` for
combinations.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _capi
from .base import Module, Parameter, positive

ActiveDims = Optional[Union[slice, Sequence[int]]]


class Kernel(Module):
    def __init__(self, active_dims: ActiveDims = None, name: Optional[str] = None):
        if active_dims is not None and not isinstance(active_dims, slice):
            active_dims = [int(i) for i in np.asarray(active_dims).reshape(-1)]
        self._active_dims = active_dims
        self.name = name or type(self).__name__.lower()

    @property
    def active_dims(self):
        return self._active_dims

    @active_dims.setter
    def active_dims(self, value):
        if value is not None and not isinstance(value, slice):
            value = [int(i) for i in np.asarray(value).reshape(-1)]
        self._active_dims = value

    def _children(self):
        # 'name' is bookkeeping, not a tf.Module child
        for key, val in super()._children():
            if key != "name":
                yield key, val

    def resolved_dims(self, D: int) -> List[int]:
        ad = self._active_dims
        if ad is None:
            return list(range(D))
        if isinstance(ad, slice):
            return list(range(D))[ad]
        for i in ad:
            if i < 0 or i >= D:
                raise ValueError(f"active_dims {ad} out of range for input dimension {D}")
        return list(ad)

    def __add__(self, other):
        return Sum([self, other])

    def __mul__(self, other):
        return Product([self, other])

    # Evaluation through the engine (convenience; models call the engine directly)
    def K(self, X, X2=None):
        from .ops import kernel_matrix
        return kernel_matrix(self, X, X2)

    def K_diag(self, X):
        from .ops import kernel_diag
        return kernel_diag(self, X)

    def __call__(self, X, X2=None, *, full_cov: bool = True):
        if not full_cov:
            if X2 is not None:
                raise ValueError("Ambiguous inputs: `not full_cov` and `X2` are not compatible.")
            return self.K_diag(X)
        return self.K(X, X2)


class Stationary(Kernel):
    _leaf_kind: int = -1

    def __init__(self, variance=1.0, lengthscales=1.0, active_dims: ActiveDims = None, name: Optional[str] = None):
        super().__init__(active_dims, name)
        self.variance = Parameter(variance, transform=positive(), name="variance")
        self.lengthscales = Parameter(lengthscales, transform=positive(), name="lengthscales")

    @property
    def ard(self) -> bool:
        return self.lengthscales.shape != ()


class SquaredExponential(Stationary):
    _leaf_kind = _capi.LEAF_SE


RBF = SquaredExponential


class RationalQuadratic(Stationary):
    _leaf_kind = _capi.LEAF_RQ

    def __init__(self, variance=1.0, lengthscales=1.0, alpha=1.0, active_dims: ActiveDims = None, name=None):
        super().__init__(variance, lengthscales, active_dims, name)
        self.alpha = Parameter(alpha, transform=positive(), name="alpha")


class Matern12(Stationary):
    _leaf_kind = _capi.LEAF_MATERN12


class Exponential(Stationary):
    _leaf_kind = _capi.LEAF_EXPONENTIAL


class Matern32(Stationary):
    _leaf_kind = _capi.LEAF_MATERN32


class Matern52(Stationary):
    _leaf_kind = _capi.LEAF_MATERN52


class Linear(Kernel):
    _leaf_kind = _capi.LEAF_LINEAR

    def __init__(self, variance=1.0, active_dims: ActiveDims = None, name: Optional[str] = None):
        super().__init__(active_dims, name)
        self.variance = Parameter(variance, transform=positive(), name="variance")
        if self.variance.shape != ():
            raise NotImplementedError("ARD Linear variance is not supported by the fused assembly")


class Periodic(Kernel):
    def __init__(self, base_kernel: Stationary, period=1.0, name: Optional[str] = None):
        if not isinstance(base_kernel, Stationary):
            raise TypeError("Periodic requires an IsotropicStationary kernel as the `base_kernel`")
        super().__init__(None, name)
        self.base_kernel = base_kernel
        self.period = Parameter(period, transform=positive(), name="period")
        if self.period.shape != ():
            raise NotImplementedError("per-dimension periods are not supported by the fused assembly")

    @property
    def active_dims(self):
        return self.base_kernel.active_dims

    @active_dims.setter
    def active_dims(self, value):
        self.base_kernel.active_dims = value

    def resolved_dims(self, D: int) -> List[int]:
        return self.base_kernel.resolved_dims(D)


class Combination(Kernel):
    def __init__(self, kernels: Sequence[Kernel], name: Optional[str] = None):
        super().__init__(None, name)
        if not all(isinstance(k, Kernel) for k in kernels):
            raise TypeError("can only combine Kernel instances")
        flat: List[Kernel] = []
        for k in kernels:  # gpflow Combination._set_kernels: same-type combinations are flattened
            if isinstance(k, type(self)):
                flat.extend(k.kernels)
            else:
                flat.append(k)
        self.kernels = flat


class Sum(Combination):
    pass


class Product(Combination):
    pass


# ---- lowering to the device descriptor -------------------------------------------------------------


class CompiledKernel:
    """A kernel expression lowered for input dimension D: the ctypes spec plus the flat list of
    Parameters whose constrained values form theta (and receive the gradient)."""

    def __init__(self, spec: _capi.GpbKernelSpec, params: List[Parameter], offsets: List[int], n_params: int, token):
        self.spec = spec
        self.params = params
        self.offsets = offsets
        self.n_params = n_params
        self.token = token

    def theta(self) -> np.ndarray:
        out = np.empty(self.n_params, dtype=np.float64)
        for p, o in zip(self.params, self.offsets):
            v = p.numpy().reshape(-1)
            out[o:o + v.size] = v
        return out

    def scatter_grad(self, g_theta: np.ndarray) -> Dict[int, np.ndarray]:
        """constrained-theta gradient -> {id(parameter): gradient w.r.t. its UNCONSTRAINED variable}."""
        out: Dict[int, np.ndarray] = {}
        for p, o in zip(self.params, self.offsets):
            u = p.unconstrained_variable._value
            gc = g_theta[o:o + u.size].reshape(u.shape)
            out[id(p)] = gc * p.transform.forward_grad(u)
        return out


def structure_token(kernel: Kernel, D: int):
    """Cheap identity of (structure, active dims, parameter shapes): recompile only when it changes."""
    if isinstance(kernel, Combination):
        return (type(kernel).__name__, tuple(structure_token(k, D) for k in kernel.kernels))
    if isinstance(kernel, Periodic):
        return ("Periodic", id(kernel), id(kernel.period), structure_token(kernel.base_kernel, D))
    ad = kernel.active_dims
    ad_t = ("slice", ad.start, ad.stop, ad.step) if isinstance(ad, slice) else (None if ad is None else tuple(ad))
    shapes = tuple((k, id(v), v.shape) for k, v in sorted(vars(kernel).items()) if isinstance(v, Parameter))
    return (type(kernel).__name__, id(kernel), ad_t, shapes, D)


def compile_kernel(kernel: Kernel, D: int) -> CompiledKernel:
    if D < 1 or D > _capi.GPB_MAX_DIMS:
        raise ValueError(f"input dimension {D} outside the supported range [1, {_capi.GPB_MAX_DIMS}]")
    params: List[Parameter] = []
    offsets: List[int] = []
    index_of: Dict[int, int] = {}
    n_params = 0

    def param_index(p: Parameter) -> int:
        nonlocal n_params
        if id(p) not in index_of:
            index_of[id(p)] = n_params
            params.append(p)
            offsets.append(n_params)
            n_params += max(1, int(np.prod(p.shape)) if p.shape else 1)
        return index_of[id(p)]

    groups: List[Tuple] = []       # (kind, mask, ard_param_id, period_param_id) -> dedupe key
    group_structs: List[_capi.GpbGroup] = []
    leaves: List[_capi.GpbLeaf] = []
    leaf_of: Dict[int, int] = {}

    def group_index(kind: int, dims: List[int], ard_param: Optional[Parameter], period_param: Optional[Parameter]) -> int:
        mask = 0
        for d in dims:
            mask |= (1 << d)
        if len(set(dims)) != len(dims):
            raise NotImplementedError("repeated columns in active_dims are not supported")
        if ard_param is not None and list(dims) != sorted(dims):
            raise NotImplementedError("ARD lengthscales need increasing active_dims")
        key = (kind, mask, id(ard_param) if ard_param is not None else None,
               id(period_param) if period_param is not None else None)
        for gi, k in enumerate(groups):
            if k == key:
                return gi
        g = _capi.GpbGroup()
        g.kind = kind
        g.dim_mask = mask
        g.ard_index = param_index(ard_param) if ard_param is not None else -1
        g.period_index = param_index(period_param) if period_param is not None else -1
        groups.append(key)
        group_structs.append(g)
        return len(groups) - 1

    def leaf_index(k: Kernel) -> int:
        if id(k) in leaf_of:
            return leaf_of[id(k)]
        lf = _capi.GpbLeaf()
        lf.alpha_index = -1
        lf.ls_index = -1
        if isinstance(k, Periodic):
            b = k.base_kernel
            dims = b.resolved_dims(D)
            r_kind = b._leaf_kind in (_capi.LEAF_MATERN12, _capi.LEAF_EXPONENTIAL, _capi.LEAF_MATERN32, _capi.LEAF_MATERN52)
            gkind = _capi.GROUP_PERIODIC_ABS if r_kind else _capi.GROUP_PERIODIC_SQ
            lf.kind = b._leaf_kind
            # parameter order follows GPflow: base_kernel.(alpha, lengthscales, variance), then period
            if isinstance(b, RationalQuadratic):
                lf.alpha_index = param_index(b.alpha)
            if b.ard:
                if b.lengthscales.shape != (len(dims),):
                    raise ValueError("ARD lengthscales must have one entry per active dimension")
                param_index(b.lengthscales)
            else:
                lf.ls_index = param_index(b.lengthscales)
            lf.var_index = param_index(b.variance)
            param_index(k.period)
            lf.group = group_index(gkind, dims, b.lengthscales if b.ard else None, k.period)
        elif isinstance(k, Linear):
            lf.kind = _capi.LEAF_LINEAR
            lf.var_index = param_index(k.variance)
            lf.group = group_index(_capi.GROUP_DOT, k.resolved_dims(D), None, None)
        elif isinstance(k, Stationary):
            dims = k.resolved_dims(D)
            lf.kind = k._leaf_kind
            if isinstance(k, RationalQuadratic):
                lf.alpha_index = param_index(k.alpha)
            if k.ard:
                if k.lengthscales.shape != (len(dims),):
                    raise ValueError("ARD lengthscales must have one entry per active dimension")
                param_index(k.lengthscales)
            else:
                lf.ls_index = param_index(k.lengthscales)
            lf.var_index = param_index(k.variance)
            lf.group = group_index(_capi.GROUP_EUCLID, dims, k.lengthscales if k.ard else None, None)
        else:
            raise TypeError(f"unsupported kernel type {type(k).__name__}")
        leaves.append(lf)
        leaf_of[id(k)] = len(leaves) - 1
        return leaf_of[id(k)]

    def sop(k: Kernel) -> List[List[int]]:
        """sum-of-products normal form: list of terms, each a list of leaf indices."""
        if isinstance(k, Sum):
            out: List[List[int]] = []
            for c in k.kernels:
                out.extend(sop(c))
            return out
        if isinstance(k, Product):
            out = [[]]
            for c in k.kernels:
                cs = sop(c)
                out = [a + b for a in out for b in cs]
            return out
        return [[leaf_index(k)]]

    terms = sop(kernel)
    if len(terms) > _capi.GPB_MAX_TERMS:
        raise NotImplementedError(f"kernel expands to {len(terms)} product terms (max {_capi.GPB_MAX_TERMS})")
    if len(leaves) > _capi.GPB_MAX_LEAVES:
        raise NotImplementedError(f"kernel has {len(leaves)} leaves (max {_capi.GPB_MAX_LEAVES})")
    if len(group_structs) > _capi.GPB_MAX_GROUPS:
        raise NotImplementedError(f"kernel needs {len(group_structs)} distance groups (max {_capi.GPB_MAX_GROUPS})")
    if n_params > _capi.GPB_MAX_PARAMS:
        raise NotImplementedError(f"kernel has {n_params} hyper-parameters (max {_capi.GPB_MAX_PARAMS})")
    spec = _capi.GpbKernelSpec()
    spec.n_dims = D
    spec.n_params = n_params
    spec.n_groups = len(group_structs)
    spec.n_leaves = len(leaves)
    spec.n_terms = len(terms)
    for i, g in enumerate(group_structs):
        spec.groups[i] = g
    for i, lf in enumerate(leaves):
        spec.leaves[i] = lf
    for i, t in enumerate(terms):
        if len(t) > _capi.GPB_MAX_FACTORS:
            raise NotImplementedError(f"product of {len(t)} leaves (max {_capi.GPB_MAX_FACTORS})")
        spec.terms[i].n_factors = len(t)
        for f, li in enumerate(t):
            spec.terms[i].leaf[f] = li
    return CompiledKernel(spec, params, offsets, n_params, (structure_token(kernel, D),))
