"""CPU stand-in for portfoliooptgp_b200._capi.Engine, for tests of the HOST layer only.

It receives exactly what the C-ABI receives -- the lowered ``gpb_kernel_spec``, the flat
constrained theta vector and raw buffer addresses -- rebuilds an oracle kernel tree from the spec
and answers with the CPU oracle.  That exercises kernel lowering, parameter flattening, the
softplus chain rule, gradient scattering, Scipy packing and the lock-step driver without a GPU.
Test infrastructure: never used by the product (which has no CPU fallback)."""
import ctypes

import numpy as np

from oracle import gpflow_oracle as O
from portfoliooptgp_b200 import _capi

_KIND = {_capi.LEAF_SE: "se", _capi.LEAF_RQ: "rq", _capi.LEAF_MATERN12: "matern12", _capi.LEAF_EXPONENTIAL: "exponential",
         _capi.LEAF_MATERN32: "matern32", _capi.LEAF_MATERN52: "matern52", _capi.LEAF_LINEAR: "linear"}


def _arr(ptr, shape):
    n = int(np.prod(shape))
    buf = (ctypes.c_double * n).from_address(int(ptr))
    return np.ctypeslib.as_array(buf).reshape(shape)


def spec_to_oracle(spec, theta):
    """gpb_kernel_spec + theta -> (oracle kernel tree, [(owner, attr, theta_index, size)])."""
    slots = []
    leaves = []
    for l in range(spec.n_leaves):
        lf = spec.leaves[l]
        g = spec.groups[lf.group]
        dims = [d for d in range(spec.n_dims) if (g.dim_mask >> d) & 1]
        kind = _KIND[lf.kind]
        if g.ard_index >= 0:
            ls = np.array(theta[g.ard_index:g.ard_index + len(dims)], dtype=np.float64)
        elif lf.ls_index >= 0:
            ls = float(theta[lf.ls_index])
        else:
            ls = 1.0
        leaf = O.Leaf(kind, variance=float(theta[lf.var_index]), lengthscales=ls,
                      alpha=float(theta[lf.alpha_index]) if lf.alpha_index >= 0 else 1.0, active_dims=dims)
        node = leaf
        if g.kind in (_capi.GROUP_PERIODIC_SQ, _capi.GROUP_PERIODIC_ABS):
            node = O.Periodic(leaf, float(theta[g.period_index]))
        leaves.append((node, leaf, lf, g, len(dims)))
    terms = []
    for t in range(spec.n_terms):
        tm = spec.terms[t]
        fac = [leaves[tm.leaf[f]][0] for f in range(tm.n_factors)]
        terms.append(fac[0] if len(fac) == 1 else O.Product(fac))
    return (terms[0] if len(terms) == 1 else O.Sum(terms)), leaves


def grad_to_theta_order(spec, leaves, kernel, g_oracle):
    """oracle gradient (O.get_theta order of the rebuilt tree) -> engine theta order.  Leaves that occur
    in several terms appear several times in the oracle tree: contributions are summed per theta index."""
    out = np.zeros(spec.n_params)
    # walk the oracle tree in O.kernel_params order and map each entry back through object identity
    index_of = {}
    for node, leaf, lf, g, nd in leaves:
        if lf.alpha_index >= 0:
            index_of[(id(leaf), "alpha")] = (lf.alpha_index, 1)
        if g.ard_index >= 0:
            index_of[(id(leaf), "lengthscales")] = (g.ard_index, nd)
        elif lf.ls_index >= 0:
            index_of[(id(leaf), "lengthscales")] = (lf.ls_index, 1)
        index_of[(id(leaf), "variance")] = (lf.var_index, 1)
        if isinstance(node, O.Periodic):
            index_of[(id(node), "period")] = (g.period_index, 1)
    pos = 0
    for _, owner, attr in O.kernel_params(kernel):
        n = int(np.size(getattr(owner, attr)))
        if (id(owner), attr) in index_of:
            i0, cnt = index_of[(id(owner), attr)]
            out[i0:i0 + cnt] += g_oracle[pos:pos + n]
        pos += n
    return out


class FakeEngine:
    device = 0

    def __init__(self):
        self.spec = None
        self.launches = 0

    def set_stream(self, s):
        pass

    def launch_count(self):
        return self.launches

    def set_kernel(self, spec, token=None):
        self.spec = spec

    def gpr_set_data(self, dX, N, D, dYc):
        self.X = _arr(dX, (N, D)).copy()
        self.Y = _arr(dYc, (N, 1)).copy()

    def _kernel(self, theta):
        return spec_to_oracle(self.spec, np.asarray(theta, dtype=np.float64))

    def gpr_lml(self, theta, noise):
        k, _ = self._kernel(theta)
        self.launches += 1
        return O.gpr_lml(k, self.X, self.Y, noise)

    def gpr_lml_grad(self, theta, noise):
        k, leaves = self._kernel(theta)
        lml, g, gn = O.gpr_lml_and_grad(k, self.X, self.Y, noise)
        self.launches += 1
        return lml, grad_to_theta_order(self.spec, leaves, k, g), gn

    def gpr_predict_f(self, theta, noise, dXs, Ns, dmean, dvar):
        k, _ = self._kernel(theta)
        Xs = _arr(dXs, (Ns, self.X.shape[1]))
        m, v = O.gpr_predict_f(k, self.X, self.Y, noise, Xs)
        _arr(dmean, (Ns,))[:] = m[:, 0]
        _arr(dvar, (Ns,))[:] = v[:, 0]

    # ---- SGPR (gpb_sgpr_elbo / gpb_sgpr_predict_f) -----------------------------------------------
    def sgpr_elbo(self, theta, noise, dZ, M, D, dX, derr, N, n_params, want_grad, derrbar=None):
        from oracle import gpflow_oracle_torch as T
        k, leaves = self._kernel(theta)
        Z, X, err = _arr(dZ, (M, D)).copy(), _arr(dX, (N, D)).copy(), _arr(derr, (N, 1)).copy()
        out = np.zeros(2 + n_params + M * D)
        self.launches += 1
        if not want_grad:
            out[0] = O.sgpr_elbo(k, Z, noise, X, err)
            return out
        e, g = T.sgpr_elbo_and_grad(k, Z, noise, X, err)
        out[0], out[1] = e, g["noise"]
        out[2:2 + n_params] = grad_to_theta_order(self.spec, leaves, k, g["theta"])
        out[2 + n_params:] = g["Z"].reshape(-1)
        if derrbar is not None:
            _arr(derrbar, (N,))[:] = g["err"]
        return out

    def sgpr_predict_f(self, theta, noise, dZ, M, D, dX, derr, N, dXs, Ns, dmean, dvar):
        k, _ = self._kernel(theta)
        Z, X, err = _arr(dZ, (M, D)).copy(), _arr(dX, (N, D)).copy(), _arr(derr, (N, 1)).copy()
        m, v = O.sgpr_predict_f(k, Z, noise, X, err, _arr(dXs, (Ns, D)))
        _arr(dmean, (Ns,))[:] = m[:, 0]
        _arr(dvar, (Ns,))[:] = v[:, 0]
