"""ctypes binding of libgpb200.so (include/gpb200.h).  The only bridge between the Python host
layer and the sm_100a engine: plain pointers and sizes, no torch types in any signature.

There is no CPU fallback: if the library is missing or no B200 is visible, construction of an
:class:`Engine` raises.  Importing this module never touches CUDA (so that the host-side logic
and the ABI checks can be tested on a CPU-only box).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgpb200.so")

GPB_MAX_DIMS = 16
GPB_MAX_GROUPS = 8
GPB_MAX_LEAVES = 8
GPB_MAX_TERMS = 8
GPB_MAX_FACTORS = 4
GPB_MAX_PARAMS = 48

GROUP_EUCLID, GROUP_PERIODIC_SQ, GROUP_PERIODIC_ABS, GROUP_DOT = 0, 1, 2, 3
LEAF_SE, LEAF_RQ, LEAF_MATERN12, LEAF_EXPONENTIAL, LEAF_MATERN32, LEAF_MATERN52, LEAF_LINEAR = range(7)


class GpbGroup(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dim_mask", C.c_uint32), ("ard_index", C.c_int32), ("period_index", C.c_int32)]


class GpbLeaf(C.Structure):
    _fields_ = [("kind", C.c_int32), ("group", C.c_int32), ("var_index", C.c_int32), ("ls_index", C.c_int32),
                ("alpha_index", C.c_int32)]


class GpbTerm(C.Structure):
    _fields_ = [("n_factors", C.c_int32), ("leaf", C.c_int32 * GPB_MAX_FACTORS)]


class GpbKernelSpec(C.Structure):
    _fields_ = [("n_dims", C.c_int32), ("n_params", C.c_int32), ("n_groups", C.c_int32), ("n_leaves", C.c_int32),
                ("n_terms", C.c_int32), ("groups", GpbGroup * GPB_MAX_GROUPS), ("leaves", GpbLeaf * GPB_MAX_LEAVES),
                ("terms", GpbTerm * GPB_MAX_TERMS)]


class CholeskyError(ValueError):
    """Raised where GPflow/TF raise InvalidArgumentError('Cholesky decomposition was not
    successful'): the covariance matrix is not positive definite at the current hyper-parameters."""


class EngineError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None

_P = C.c_void_p
_D = C.c_double
_I64 = C.c_int64
_INT = C.c_int
_DP = C.POINTER(C.c_double)

# name -> (restype, argtypes); must list every symbol include/gpb200.h declares
SIGNATURES = {
    "gpb_version": (_INT, []),
    "gpb_create": (_INT, [C.POINTER(_P), _INT]),
    "gpb_destroy": (_INT, [_P]),
    "gpb_last_error": (C.c_char_p, [_P]),
    "gpb_set_stream": (_INT, [_P, _P]),
    "gpb_launch_count": (_I64, [_P]),
    "gpb_kernel_shape": (_INT, [_P]),
    "gpb_set_option": (_INT, [_P, _INT, _INT]),
    "gpb_profile_enable": (_INT, [_P, _INT]),
    "gpb_profile_read": (_INT, [_P, _DP, C.POINTER(C.c_int64)]),
    "gpb_profile_read_flops": (_INT, [_P, _DP]),
    "gpb_set_kernel": (_INT, [_P, C.POINTER(GpbKernelSpec)]),
    "gpb_assemble": (_INT, [_P, _DP, _P, _I64, _P, _I64, _INT, _P, _I64, _INT, _D]),
    "gpb_kdiag": (_INT, [_P, _DP, _P, _I64, _INT, _P]),
    "gpb_potrf": (_INT, [_P, _P, _I64, _I64]),
    "gpb_potrf_inv": (_INT, [_P, _P, _I64, _I64, _P, _I64]),
    "gpb_lauum": (_INT, [_P, _P, _I64, _I64, _P, _I64]),
    "gpb_gemm": (_INT, [_P, _INT, _INT, _I64, _I64, _I64, _D, _P, _I64, _P, _I64, _D, _P, _I64, _INT]),
    "gpb_gpr_set_data": (_INT, [_P, _P, _I64, _INT, _P]),
    "gpb_gpr_lml": (_INT, [_P, _DP, _D, _DP]),
    "gpb_gpr_lml_grad": (_INT, [_P, _DP, _D, _DP, _DP, _DP]),
    "gpb_gpr_predict_f": (_INT, [_P, _DP, _D, _P, _I64, _P, _P]),
    "gpb_gpr_predict_f_many": (_INT, [_INT, C.POINTER(_P), _I64, C.POINTER(_P), C.POINTER(_I64), _INT, C.POINTER(_P), _DP, _INT,
                                      _DP, C.POINTER(_P), _I64, C.POINTER(_P), C.POINTER(_P), C.POINTER(_INT)]),
    "gpb_gpr_lml_grad_many": (_INT, [_INT, C.POINTER(_P), _I64, C.POINTER(_P), C.POINTER(_I64), _INT, C.POINTER(_P), _DP, _INT,
                                     _DP, _INT, _DP, _DP, _DP, C.POINTER(_INT)]),
    "gpb_gpr_get_alpha": (_INT, [_P, _P]),
    "gpb_batched_lml_grad": (_INT, [_P, _P, _P, _P, _P, _I64, _I64, _INT, _P, _P, _INT]),
    "gpb_batched_predict_f": (_INT, [_P, _P, _P, _P, _P, _I64, _I64, _INT, _P, _I64, _P, _P, _P]),
    "gpb_batched_lml_grad_ragged": (_INT, [_P, _P, _P, _P, _P, _P, _I64, _I64, _INT, _P, _P, _INT]),
    "gpb_batched_predict_f_ragged": (_INT, [_P, _P, _P, _P, _P, _P, _I64, _I64, _INT, _P, _I64, _P, _P, _P]),
    "gpb_svgp_flat_size": (_I64, [_I64, _INT, _INT]),
    "gpb_svgp_data_term": (_INT, [_P, _DP, _D, _P, _I64, _INT, _P, _P, _I64, _P, _P, _I64, _P, _INT]),
    "gpb_svgp_finish": (_INT, [_P, _P, _D, _P, _P, _I64, _I64, _INT, _INT, _INT, _DP, _DP]),
    "gpb_svgp_predict_f": (_INT, [_P, _DP, _P, _I64, _INT, _P, _P, _I64, _P, _I64, _P, _P]),
    "gpb_gpr_factor_serial": (_I64, [_P]),
    "gpb_gpr_predict_f_reuse": (_INT, [_P, _DP, _D, _I64, _P, _I64, _P, _P]),
    "gpb_sgpr_elbo": (_INT, [_P, _DP, _D, _P, _I64, _INT, _P, _P, _I64, _INT, _DP, _P]),
    "gpb_sgpr_predict_f": (_INT, [_P, _DP, _D, _P, _I64, _INT, _P, _P, _I64, _P, _I64, _P, _P]),
    "gpb_adam_step": (_INT, [_P, _P, _P, _P, _P, _I64, _D, _D, _D, _D, _I64, _INT]),
    "gpb_prep_returns": (_INT, [_P, _P, _P, _I64, _I64, _INT, _P]),
    "gpb_prep_zscore": (_INT, [_P, _P, _I64, _I64, _INT, _P, _I64, _P, _P]),
    "gpb_prep_windows": (_INT, [_P, _P, _P, _I64, _I64, _INT, _I64, _I64, _P, _P]),
    "gpb_post_upsample": (_INT, [_P, _P, _I64, _P, _I64, _P, _INT, _P]),
    "gpb_post_blend": (_INT, [_P, _D, _D, _P, _P, _P, _I64, _P]),
}


def load_library() -> C.CDLL:
    """dlopen libgpb200.so and declare every entry point.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # a fresh clone has no built library (it is git-ignored): build it in-tree once, if nvcc is here
        # (GPB_NO_AUTOBUILD=1 switches this off); there is no CPU fallback on this path either way
        err = None
        if os.environ.get("GPB_NO_AUTOBUILD", "0") in ("", "0"):
            try:
                from . import build as _build
                _build.build()
            except Exception as e:  # nvcc missing or a compile error: report it below
                err = e
        if not os.path.exists(LIB_PATH):
            raise EngineError(
                f"{LIB_PATH} is missing and could not be built ({err}): build it with "
                "`python -m portfoliooptgp_b200.build` (nvcc, sm_100a).  There is no CPU fallback on this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _as_dp(a: np.ndarray):
    return a.ctypes.data_as(_DP)


class Engine:
    """One gpb_handle bound to one CUDA device.  Device buffers are passed as integer addresses
    (``tensor.data_ptr()``); ownership stays with the caller."""

    def __init__(self, device: int = 0):
        self._h = _P()
        self._lib = load_library()
        rc = self._lib.gpb_create(C.byref(self._h), int(device))
        if rc != 0:
            self._h = _P()
            reasons = {-10: "no CUDA device visible", -13: "device is not sm_100 (B200) class", -2: "bad device index"}
            raise EngineError(f"gpb_create failed ({rc}): {reasons.get(rc, 'CUDA initialisation error')}; "
                              "the engine has no CPU fallback")
        self.device = int(device)
        self._spec_token = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.gpb_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -----------------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc == 0:
            return
        msg = self._lib.gpb_last_error(self._h)
        msg = msg.decode() if msg else ""
        if rc > 0:
            raise CholeskyError(f"{what}: {msg}")
        raise EngineError(f"{what} failed ({rc}): {msg}")

    def set_stream(self, cuda_stream: int):
        self._check(self._lib.gpb_set_stream(self._h, _P(cuda_stream)), "gpb_set_stream")

    def kernel_shape(self) -> int:
        """1-based id of the straight-line shape (csrc/shapes.cuh) the current expression matches; 0 = interpreter."""
        rc = int(self._lib.gpb_kernel_shape(self._h))
        if rc < 0:
            self._check(rc, "gpb_kernel_shape")
        return rc

    def launch_count(self) -> int:
        return int(self._lib.gpb_launch_count(self._h))

    PROF_CATEGORIES = ("gemm", "assemble", "leaf", "grad_reduce", "vector", "batched", "svgp", "gemm_small")

    OPTION_FORK_STREAMS = 0
    OPTION_PDL = 1
    OPTION_STATIC_SHAPES = 2
    OPTION_REFINE_OBJECTIVE = 3   # 0 never, 1 automatic (ill-conditioned K only), 2 always
    OPTION_CHAIN = 5              # look-ahead chain at the bottom of the blocked factorisation (default on)
    OPTION_PIPELINE = 4           # two-partition pipelined factorisation for N >= 3072 (default off: measured slower)

    def set_option(self, option: int, value: int):
        self._check(self._lib.gpb_set_option(self._h, int(option), int(value)), "gpb_set_option")

    def profile_enable(self, on: bool):
        self._check(self._lib.gpb_profile_enable(self._h, int(bool(on))), "gpb_profile_enable")

    def profile_read(self):
        ms = np.zeros(8, dtype=np.float64)
        cnt = np.zeros(8, dtype=np.int64)
        self._check(self._lib.gpb_profile_read(self._h, _as_dp(ms), cnt.ctypes.data_as(C.POINTER(C.c_int64))),
                    "gpb_profile_read")
        return ({c: float(m) for c, m in zip(self.PROF_CATEGORIES, ms)}, {c: int(n) for c, n in zip(self.PROF_CATEGORIES, cnt)})

    def profile_flops(self):
        fl = np.zeros(8, dtype=np.float64)
        self._check(self._lib.gpb_profile_read_flops(self._h, _as_dp(fl)), "gpb_profile_read_flops")
        return {c: float(f) for c, f in zip(self.PROF_CATEGORIES, fl)}

    def set_kernel(self, spec: GpbKernelSpec, token=None):
        if token is not None and token == self._spec_token:
            return
        self._check(self._lib.gpb_set_kernel(self._h, C.byref(spec)), "gpb_set_kernel")
        self._spec_token = token

    # -- assembly ----------------------------------------------------------------------------
    def assemble(self, theta: np.ndarray, dX: int, N: int, dX2: Optional[int], N2: int, D: int, dK: int, ldk: int,
                 mode: int, diag_add: float):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        self._check(self._lib.gpb_assemble(self._h, _as_dp(theta), _P(dX), N, _P(dX2) if dX2 else None, N2, D, _P(dK),
                                           ldk, mode, float(diag_add)), "gpb_assemble")

    def kdiag(self, theta: np.ndarray, dX: int, N: int, D: int, dout: int):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        self._check(self._lib.gpb_kdiag(self._h, _as_dp(theta), _P(dX), N, D, _P(dout)), "gpb_kdiag")

    # -- dense --------------------------------------------------------------------------------
    def potrf(self, dA: int, N: int, lda: int):
        self._check(self._lib.gpb_potrf(self._h, _P(dA), N, lda), "gpb_potrf")

    def potrf_inv(self, dA: int, N: int, lda: int, dW: int, ldw: int):
        self._check(self._lib.gpb_potrf_inv(self._h, _P(dA), N, lda, _P(dW), ldw), "gpb_potrf_inv")

    def lauum(self, dW: int, N: int, ldw: int, dOut: int, ldo: int):
        self._check(self._lib.gpb_lauum(self._h, _P(dW), N, ldw, _P(dOut), ldo), "gpb_lauum")

    def gemm(self, transa, transb, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc, tri=0):
        self._check(self._lib.gpb_gemm(self._h, int(transa), int(transb), M, N, K, float(alpha), _P(dA), lda, _P(dB),
                                       ldb, float(beta), _P(dC), ldc, int(tri)), "gpb_gemm")

    # -- exact GP -----------------------------------------------------------------------------
    def gpr_set_data(self, dX: int, N: int, D: int, dYc: int):
        self._check(self._lib.gpb_gpr_set_data(self._h, _P(dX), N, D, _P(dYc)), "gpb_gpr_set_data")

    def gpr_lml(self, theta: np.ndarray, noise: float) -> float:
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        out = C.c_double()
        self._check(self._lib.gpb_gpr_lml(self._h, _as_dp(theta), float(noise), C.byref(out)), "gpb_gpr_lml")
        return out.value

    def gpr_lml_grad(self, theta: np.ndarray, noise: float):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        lml = C.c_double()
        gnoise = C.c_double()
        g = np.zeros(theta.size, dtype=np.float64)
        self._check(self._lib.gpb_gpr_lml_grad(self._h, _as_dp(theta), float(noise), C.byref(lml), _as_dp(g),
                                               C.byref(gnoise)), "gpb_gpr_lml_grad")
        return lml.value, g, gnoise.value

    def gpr_predict_f(self, theta: np.ndarray, noise: float, dXs: int, Ns: int, dmean: int, dvar: int):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        self._check(self._lib.gpb_gpr_predict_f(self._h, _as_dp(theta), float(noise), _P(dXs), Ns, _P(dmean), _P(dvar)),
                    "gpb_gpr_predict_f")

    @staticmethod
    def gpr_lml_grad_many(engines, dX: np.ndarray, N: np.ndarray, D: int, dY: np.ndarray, theta: np.ndarray, noise: np.ndarray,
                          want_grad: bool = True):
        """``gpb_gpr_lml_grad_many``: job j = (X pointer dX[j], rows N[j], Y pointer dY[j], theta[j], noise[j]) on
        engine j % len(engines), one host thread of the library per engine.  Returns (lml [J], grad_theta [J,P],
        grad_noise [J], rc [J]) -- rc > 0: first non-positive pivot of that job; a negative rc raises."""
        lib = engines[0]._lib
        nh = len(engines)
        hs = (_P * nh)(*[e._h for e in engines])
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        J, P = theta.shape
        if noise.shape != (J,) or len(dX) != J or len(dY) != J or len(N) != J:
            raise ValueError("gpr_lml_grad_many: one X, Y, N, theta row and noise per job")
        xs = (_P * J)(*[int(v) for v in dX])
        ys = (_P * J)(*[int(v) for v in dY])
        ns = np.ascontiguousarray(N, dtype=np.int64)
        lml = np.zeros(J)
        g = np.zeros((J, P))
        gn = np.zeros(J)
        rc = np.zeros(J, dtype=np.int32)
        ret = lib.gpb_gpr_lml_grad_many(nh, hs, J, xs, ns.ctypes.data_as(C.POINTER(_I64)), int(D), ys, _as_dp(theta), int(P),
                                        _as_dp(noise), int(bool(want_grad)), _as_dp(lml), _as_dp(g), _as_dp(gn),
                                        rc.ctypes.data_as(C.POINTER(_INT)))
        if ret != 0:
            raise EngineError(f"gpb_gpr_lml_grad_many: bad arguments or thread creation failed ({ret})")
        neg = np.nonzero(rc < 0)[0]
        if neg.size:
            j = int(neg[0])
            msg = lib.gpb_last_error(engines[j % nh]._h)
            raise EngineError(f"gpb_gpr_lml_grad_many: job {j} failed ({int(rc[j])}): {msg.decode() if msg else ''}")
        return lml, g, gn, rc

    @staticmethod
    def gpr_predict_f_many(engines, dX: np.ndarray, N: np.ndarray, D: int, dY: np.ndarray, theta: np.ndarray, noise: np.ndarray,
                           dXs: np.ndarray, Ns: int, dmean: np.ndarray, dvar: np.ndarray) -> np.ndarray:
        """``gpb_gpr_predict_f_many``: per-job predict_f side by side; returns rc [J] (> 0: that job's covariance is
        not positive definite, its outputs are undefined); a negative rc raises.  Synchronous."""
        lib = engines[0]._lib
        nh = len(engines)
        hs = (_P * nh)(*[e._h for e in engines])
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        J, P = theta.shape
        arr = lambda v: (_P * J)(*[int(x) for x in v])
        ns = np.ascontiguousarray(N, dtype=np.int64)
        rc = np.zeros(J, dtype=np.int32)
        ret = lib.gpb_gpr_predict_f_many(nh, hs, J, arr(dX), ns.ctypes.data_as(C.POINTER(_I64)), int(D), arr(dY), _as_dp(theta), int(P),
                                         _as_dp(noise), arr(dXs), int(Ns), arr(dmean), arr(dvar), rc.ctypes.data_as(C.POINTER(_INT)))
        if ret != 0:
            raise EngineError(f"gpb_gpr_predict_f_many: bad arguments or thread creation failed ({ret})")
        neg = np.nonzero(rc < 0)[0]
        if neg.size:
            j = int(neg[0])
            msg = lib.gpb_last_error(engines[j % nh]._h)
            raise EngineError(f"gpb_gpr_predict_f_many: job {j} failed ({int(rc[j])}): {msg.decode() if msg else ''}")
        return rc

    # -- batched small GPs -----------------------------------------------------------------------
    def batched_lml_grad(self, dX: int, dYc: int, dtheta: int, dnoise: int, B: int, N: int, D: int, dout: int,
                         dinfo: int, want_grad: bool = True, dnrows: Optional[int] = None):
        if dnrows is None:
            self._check(self._lib.gpb_batched_lml_grad(self._h, _P(dX), _P(dYc), _P(dtheta), _P(dnoise), B, N, D, _P(dout),
                                                       _P(dinfo), int(bool(want_grad))), "gpb_batched_lml_grad")
        else:
            self._check(self._lib.gpb_batched_lml_grad_ragged(self._h, _P(dX), _P(dYc), _P(dtheta), _P(dnoise), _P(dnrows), B, N,
                                                              D, _P(dout), _P(dinfo), int(bool(want_grad))),
                        "gpb_batched_lml_grad_ragged")

    def batched_predict_f(self, dX: int, dYc: int, dtheta: int, dnoise: int, B: int, N: int, D: int, dXs: int, Ns: int,
                          dmean: int, dvar: int, dinfo: int, dnrows: Optional[int] = None):
        if dnrows is None:
            self._check(self._lib.gpb_batched_predict_f(self._h, _P(dX), _P(dYc), _P(dtheta), _P(dnoise), B, N, D, _P(dXs),
                                                        Ns, _P(dmean), _P(dvar), _P(dinfo)), "gpb_batched_predict_f")
        else:
            self._check(self._lib.gpb_batched_predict_f_ragged(self._h, _P(dX), _P(dYc), _P(dtheta), _P(dnoise), _P(dnrows), B, N,
                                                               D, _P(dXs), Ns, _P(dmean), _P(dvar), _P(dinfo)),
                        "gpb_batched_predict_f_ragged")

    # -- SVGP ---------------------------------------------------------------------------------------
    def svgp_flat_size(self, M: int, D: int, P: int) -> int:
        return int(self._lib.gpb_svgp_flat_size(M, D, P))

    def svgp_data_term(self, theta: np.ndarray, noise: float, dZ: int, M: int, D: int, dqmu: int, dqsqrt: int, ldq: int,
                       dXb: int, dYb: int, B: int, dflat: int, want_grad: bool = True):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        self._check(self._lib.gpb_svgp_data_term(self._h, _as_dp(theta), float(noise), _P(dZ), M, D, _P(dqmu), _P(dqsqrt),
                                                 ldq, _P(dXb), _P(dYb), B, _P(dflat), int(bool(want_grad))),
                    "gpb_svgp_data_term")

    def svgp_finish(self, dflat: int, scale: float, dqmu: int, dqsqrt: int, ldq: int, M: int, D: int, P: int,
                    apply_grad: bool = True):
        elbo, kl = C.c_double(), C.c_double()
        self._check(self._lib.gpb_svgp_finish(self._h, _P(dflat), float(scale), _P(dqmu), _P(dqsqrt), ldq, M, D, P,
                                              int(bool(apply_grad)), C.byref(elbo), C.byref(kl)), "gpb_svgp_finish")
        return elbo.value, kl.value

    def svgp_predict_f(self, theta: np.ndarray, dZ: int, M: int, D: int, dqmu: int, dqsqrt: int, ldq: int, dXs: int,
                       Ns: int, dmean: int, dvar: int):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        self._check(self._lib.gpb_svgp_predict_f(self._h, _as_dp(theta), _P(dZ), M, D, _P(dqmu), _P(dqsqrt), ldq, _P(dXs),
                                                 Ns, _P(dmean), _P(dvar)), "gpb_svgp_predict_f")

    def sgpr_elbo(self, theta, noise: float, dZ: int, M: int, D: int, dX: int, derr: int, N: int, n_params: int,
                  want_grad: bool, derrbar: Optional[int] = None) -> np.ndarray:
        """[elbo, d/dnoise, d/dtheta (n_params), d/dZ (M*D)]; only [0] is meaningful without want_grad."""
        out = np.zeros(2 + n_params + M * D, dtype=np.float64)
        self._check(self._lib.gpb_sgpr_elbo(self._h, _as_dp(theta), float(noise), _P(dZ), M, D, _P(dX), _P(derr), N,
                                            int(bool(want_grad)), out.ctypes.data_as(_DP), _P(derrbar)), "gpb_sgpr_elbo")
        return out

    def sgpr_predict_f(self, theta, noise: float, dZ: int, M: int, D: int, dX: int, derr: int, N: int, dXs: int, Ns: int,
                       dmean: int, dvar: int):
        self._check(self._lib.gpb_sgpr_predict_f(self._h, _as_dp(theta), float(noise), _P(dZ), M, D, _P(dX), _P(derr), N,
                                                 _P(dXs), Ns, _P(dmean), _P(dvar)), "gpb_sgpr_predict_f")

    def adam_step(self, dx: int, dg: int, dm: int, dv: int, n: int, lr: float, step: int, beta1=0.9, beta2=0.999,
                  eps=1e-8, maximize: bool = True):
        self._check(self._lib.gpb_adam_step(self._h, _P(dx), _P(dg), _P(dm), _P(dv), n, float(lr), float(beta1),
                                            float(beta2), float(eps), int(step), int(bool(maximize))), "gpb_adam_step")

    def prep_returns(self, dclose: int, dopen: int, T: int, A: int, kind: int, dout: int):
        self._check(self._lib.gpb_prep_returns(self._h, _P(dclose), _P(dopen), T, A, int(kind), _P(dout)),
                    "gpb_prep_returns")

    def prep_zscore(self, dx: int, T: int, A: int, ddof: int, dout: int, ldo: int, dmean: int, dstd: int):
        self._check(self._lib.gpb_prep_zscore(self._h, _P(dx), T, A, int(ddof), _P(dout), ldo, _P(dmean), _P(dstd)),
                    "gpb_prep_zscore")

    def prep_windows(self, dfeat: int, dy: int, S: int, T: int, D: int, N: int, stride: int, dX: int, dY: int):
        self._check(self._lib.gpb_prep_windows(self._h, _P(dfeat), _P(dy), S, T, D, N, stride, _P(dX), _P(dY)),
                    "gpb_prep_windows")

    def post_upsample(self, dXd: int, Nd: int, dXs: int, Ns: int, dpred: int, Q: int, dout: int):
        self._check(self._lib.gpb_post_upsample(self._h, _P(dXd), Nd, _P(dXs), Ns, _P(dpred), Q, _P(dout)), "gpb_post_upsample")

    def post_blend(self, alpha: float, beta: float, dd: int, dw: int, dm: int, n: int, dout: int):
        self._check(self._lib.gpb_post_blend(self._h, float(alpha), float(beta), _P(dd), _P(dw), _P(dm), n, _P(dout)),
                    "gpb_post_blend")

    def gpr_factor_serial(self) -> int:
        return int(self._lib.gpb_gpr_factor_serial(self._h))

    def gpr_predict_f_reuse(self, theta, noise: float, serial: int, dXs: int, Ns: int, dmean: int, dvar: int):
        self._check(self._lib.gpb_gpr_predict_f_reuse(self._h, _as_dp(theta), float(noise), int(serial), _P(dXs), Ns,
                                                      _P(dmean), _P(dvar)), "gpb_gpr_predict_f_reuse")

    def gpr_get_alpha(self, dalpha: int):
        self._check(self._lib.gpb_gpr_get_alpha(self._h, _P(dalpha)), "gpb_gpr_get_alpha")
