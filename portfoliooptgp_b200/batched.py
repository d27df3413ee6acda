"""Batched independent small GPs (north_star subsystem 4; BASELINE config C3).

The reference fits one small GPR per (asset, rolling window, restart) in nested Python loops
(Multi-Input_GPR/main.py:414-456 calling models/model_trainer.py:17-54): fresh
``gpflow.models.GPR((X[:i], Y[:i]), deepcopy(kernel), noise_variance=...)`` + ``Scipy().minimize``
+ ``predict_f(...)[-1]`` each time.  ``BatchedGPR`` holds B such problems that share one kernel
*expression* (each with its own hyper-parameter values) and evaluates all objectives and
gradients in one launch (one GP per CTA, covariance resident in shared memory), and
``lockstep_lbfgsb`` advances B SciPy L-BFGS-B instances together -- same compiled ``setulb``
reverse-communication loop as ``scipy.optimize.minimize(method="L-BFGS-B")``, so the iterates of
every problem are SciPy's own -- with one batched objective launch per round.

Across GPUs the batch is split into contiguous blocks, one per rank, with no communication during
evaluation and one final gather (``shard_range`` / ``gather_results``).
"""
from __future__ import annotations

import os

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import scipy.optimize
import torch

from . import _capi, ops
from .base import Softplus
from .kernels import Kernel, compile_kernel
from .likelihoods import DEFAULT_VARIANCE_LOWER_BOUND


# ---- lock-step L-BFGS-B ------------------------------------------------------------------------------


_SCIPY_CHECKED = None


def _scipy_lbfgsb_module():
    """The lock-step driver re-drives SciPy's PRIVATE reverse-communication routine
    (``scipy.optimize._lbfgsb_py._lbfgsb.setulb``, the C translation shipped since SciPy 1.15: integer
    task / ln_task arrays, no csave / iprint).  Checked once; an older or changed SciPy gets a clear error
    instead of a TypeError from the first fit (tested range: SciPy 1.15 - 1.18)."""
    global _SCIPY_CHECKED
    if _SCIPY_CHECKED is None:
        from scipy.optimize import _lbfgsb_py as _lb
        ver = tuple(int(v) for v in scipy.__version__.split(".")[:2] if v.isdigit())
        missing = [a for a in ("_lbfgsb", "status_messages", "task_messages") if not hasattr(_lb, a)]
        if not missing and not hasattr(_lb._lbfgsb, "setulb"):
            missing.append("_lbfgsb.setulb")
        if missing or ver < (1, 15):
            raise RuntimeError(
                f"portfoliooptgp_b200.batched needs SciPy >= 1.15 with the C L-BFGS-B driver (found SciPy {scipy.__version__}"
                f"{', missing ' + ', '.join(missing) if missing else ''}); fit the GPs one by one with optimizers.Scipy instead")
        _SCIPY_CHECKED = _lb
    return _SCIPY_CHECKED


def lockstep_lbfgsb(fun_batch: Callable[[np.ndarray, np.ndarray], Tuple[np.ndarray, np.ndarray]], X0: np.ndarray,
                    maxiter: int = 15000, maxfun: int = 15000, maxcor: int = 10, ftol: float = 2.2204460492503131e-09,
                    gtol: float = 1e-5, maxls: int = 20, workers: int = 0,
                    fun_batch_async: Optional[Callable] = None) -> List[scipy.optimize.OptimizeResult]:
    """Minimise B independent problems with SciPy's L-BFGS-B, advancing them in lock step.

    ``fun_batch(X [b, n], idx [b]) -> (f [b], g [b, n])`` evaluates the problems ``idx`` at ``X``.
    Each problem runs the reverse-communication loop of ``scipy.optimize._lbfgsb_py._minimize_lbfgsb``
    (same ``_lbfgsb.setulb``, same defaults, same stopping rules); whenever a problem asks for
    f and g it is parked until every active problem has asked, then one ``fun_batch`` call serves
    them all.

    ``fun_batch_async(X, idx) -> wait`` (optional; ``wait() -> (f, g)``) starts an evaluation without
    blocking.  With it the problems are split into two halves that alternate: while the device evaluates
    one half the host advances the other half's SciPy state machines, so device time, copies and launch
    latency disappear behind the host loop.  The sequence of (x, f, g) every problem sees is unchanged, so
    the iterates stay bit-identical to SciPy's."""
    _lb = _scipy_lbfgsb_module()
    if workers and workers > 1 and len(X0) >= 4 * workers:
        return _lockstep_lbfgsb_workers(fun_batch, X0, maxiter, maxfun, maxcor, ftol, gtol, maxls, int(workers), fun_batch_async)
    _lbfgsb = _lb._lbfgsb
    int_dtype = np.int64 if getattr(_lb, "HAS_ILP64", False) else np.int32
    X0 = np.ascontiguousarray(X0, dtype=np.float64)
    B, n = X0.shape
    m = maxcor
    factr = ftol / np.finfo(float).eps

    # State of all problems in 2-D arrays; setulb gets ROW VIEWS of them (contiguous), so that the only
    # per-problem Python work in a round is the setulb call itself and scattering / gathering x, f, g is
    # vectorised.  Profile (5120 problems, n = 5): with a multi-threaded BLAS setulb takes ~13 us per
    # call (thread hand-off on 20 x 20 matrices), with one BLAS thread ~3 us -- hence the
    # threadpool_limits(1) around the loop below (same arithmetic, same iterates).  SciPy's routine is kept
    # because it reproduces scipy.optimize's iterates bit for bit; ``workers`` > 1
    # spreads the calls over processes for another 2-3x.
    X = np.array(X0, dtype=np.float64)
    F = np.zeros((B,), dtype=np.float64)
    G = np.zeros((B, n), dtype=np.float64)
    WA = np.zeros((B, 2 * m * n + 5 * n + 11 * m * m + 8 * m), dtype=np.float64)
    IWA = np.zeros((B, 3 * n), dtype=int_dtype)
    TASK = np.zeros((B, 2), dtype=int_dtype)
    LN_TASK = np.zeros((B, 2), dtype=int_dtype)
    LSAVE = np.zeros((B, 4), dtype=int_dtype)
    ISAVE = np.zeros((B, 44), dtype=int_dtype)
    DSAVE = np.zeros((B, 29), dtype=np.float64)
    NIT = np.zeros((B,), dtype=np.int64)
    NFEV = np.zeros((B,), dtype=np.int64)
    nbd = np.zeros(n, dtype=int_dtype)
    low = np.zeros(n, dtype=np.float64)
    up = np.zeros(n, dtype=np.float64)
    args = [(m, X[b], low, up, nbd, F[b:b + 1].reshape(()), G[b], factr, gtol, WA[b], IWA[b], TASK[b], LSAVE[b], ISAVE[b],
             DSAVE[b], maxls, LN_TASK[b]) for b in range(B)]
    setulb = _lbfgsb.setulb
    try:
        from threadpoolctl import threadpool_limits
        blas_single = threadpool_limits(limits=1)
    except Exception:  # pragma: no cover - threadpoolctl ships with scikit-learn; the loop works without it
        blas_single = None

    def advance(group):
        """run every problem of ``group`` until it needs f, g (returned) or stops"""
        waiting = []
        for b in group:
            a = args[b]
            task = a[11]
            while True:
                setulb(*a)
                t0 = task[0]
                if t0 == 3:
                    waiting.append(b)
                    break
                elif t0 == 1:
                    NIT[b] += 1
                    if NIT[b] >= maxiter:
                        task[0] = 5
                        task[1] = 504
                    elif NFEV[b] > maxfun:
                        task[0] = 5
                        task[1] = 502
                else:
                    break
        return waiting

    def scatter(idx, fb, gb):
        F[idx] = np.asarray(fb, dtype=np.float64)
        G[idx] = np.asarray(gb, dtype=np.float64)
        NFEV[idx] += 1

    try:
        if fun_batch_async is not None and B >= 2:
            # two halves in flight: advance one on the host while the device evaluates the other
            groups = [list(range(0, B // 2)), list(range(B // 2, B))]
            pending = [None, None]
            for h in (0, 1):
                groups[h] = advance(groups[h])
                if groups[h]:
                    idx = np.asarray(groups[h], dtype=np.int64)
                    pending[h] = (idx, fun_batch_async(X[idx], idx))
            h = 0
            while pending[0] is not None or pending[1] is not None:
                if pending[h] is not None:
                    idx, wait = pending[h]
                    fb, gb = wait()
                    scatter(idx, fb, gb)
                    groups[h] = advance(groups[h])
                    pending[h] = None
                    if groups[h]:
                        idx = np.asarray(groups[h], dtype=np.int64)
                        pending[h] = (idx, fun_batch_async(X[idx], idx))
                h ^= 1
        else:
            active = list(range(B))
            while active:
                waiting = advance(active)
                if not waiting:
                    break
                idx = np.asarray(waiting, dtype=np.int64)
                fb, gb = fun_batch(X[idx], idx)
                scatter(idx, fb, gb)
                active = waiting
    finally:
        if blas_single is not None:
            blas_single.restore_original_limits()
    results = []
    for b in range(B):
        t0, t1 = int(TASK[b, 0]), int(TASK[b, 1])
        nit, nfev = int(NIT[b]), int(NFEV[b])
        if t0 == 4:
            warnflag = 0
        elif nfev > maxfun or nit >= maxiter:
            warnflag = 1
        else:
            warnflag = 2
        msg = _lb.status_messages[t0] + ": " + _lb.task_messages[t1]
        results.append(scipy.optimize.OptimizeResult(fun=float(F[b]), jac=G[b].copy(), nfev=nfev, njev=nfev, nit=nit,
                                                     status=warnflag, message=msg, x=X[b].copy(), success=(warnflag == 0)))
    return results


def _lockstep_lbfgsb_workers(fun_batch, X0, maxiter, maxfun, maxcor, ftol, gtol, maxls, workers, fun_batch_async=None):
    """lockstep_lbfgsb with the per-problem setulb calls spread over worker processes (_lbfgsb_pool):
    the same call sequence per problem, so the same iterates; only the host time per round shrinks.
    With ``fun_batch_async`` and >= 4 workers the workers form two groups that alternate: while one group's
    problems are evaluated on the device the other group's processes advance their SciPy state machines."""
    _lb = _scipy_lbfgsb_module()
    from . import _lbfgsb_pool
    X0 = np.ascontiguousarray(X0, dtype=np.float64)
    B, n = X0.shape
    factr = ftol / np.finfo(float).eps
    bounds = [(B * w) // workers for w in range(workers + 1)]
    pool = _lbfgsb_pool.get_workers(workers)
    for w, wk in enumerate(pool):
        wk.send(("init", X0[bounds[w]:bounds[w + 1]], maxcor, factr, gtol, maxls, maxiter, maxfun))
    replies = [wk.recv() for wk in pool]
    pipelined = fun_batch_async is not None and workers >= 4
    ngroups = (4 if workers >= 12 else 3 if workers >= 9 else 2) if pipelined else 1
    if pipelined and os.environ.get("GPB_FIT_GROUPS"):       # tuning knob (tools/c3_fit.py)
        ngroups = max(2, min(workers // 2, int(os.environ["GPB_FIT_GROUPS"])))
    cuts = [(workers * g) // ngroups for g in range(ngroups + 1)]
    groups = [list(range(cuts[g], cuts[g + 1])) for g in range(ngroups)]

    def gather(g):
        counts = {w: len(replies[w][0]) for w in g}
        if sum(counts.values()) == 0:
            return None
        idx = np.concatenate([replies[w][0] + bounds[w] for w in g])
        Xb = np.concatenate([replies[w][1] for w in g], axis=0)
        return counts, idx, Xb

    def scatter(g, counts, fb, gb):
        fb = np.asarray(fb, dtype=np.float64)
        gb = np.asarray(gb, dtype=np.float64)
        o = 0
        for w in g:
            c = counts[w]
            if c:
                pool[w].send(("step", fb[o:o + c], gb[o:o + c]))
            o += c

    if not pipelined:
        while True:
            got = gather(groups[0])
            if got is None:
                break
            counts, idx, Xb = got
            fb, gb = fun_batch(Xb, idx)
            scatter(groups[0], counts, fb, gb)
            for w in groups[0]:
                if counts[w]:
                    replies[w] = pool[w].recv()
    else:
        # Event-driven: every group is either on the device (an evaluation in flight) or on the host (its
        # processes advancing); the parent serves whichever is ready, so one group's device time hides behind
        # the other groups' host time and nobody waits at the head of a fixed order.
        import select
        EVAL, ADV, DONE = 0, 1, 2
        state = [DONE] * ngroups
        info = [None] * ngroups

        def launch(gi):
            got = gather(groups[gi])
            if got is None:
                state[gi], info[gi] = DONE, None
            else:
                state[gi], info[gi] = EVAL, (got[0], fun_batch_async(got[2], got[1]))

        for gi in range(ngroups):
            launch(gi)
        while any(st != DONE for st in state):
            progressed = False
            for gi in range(ngroups):
                if state[gi] == EVAL:
                    counts, wait = info[gi]
                    if getattr(wait, "ready", None) is None or wait.ready():
                        fb, gb = wait()
                        scatter(groups[gi], counts, fb, gb)
                        state[gi], info[gi] = ADV, {w for w in groups[gi] if counts[w]}
                        progressed = True
                elif state[gi] == ADV:
                    waiting = info[gi]
                    ready, _, _ = select.select([pool[w] for w in waiting], [], [], 0)
                    for wk in ready:
                        w = pool.index(wk)
                        replies[w] = wk.recv()
                        waiting.discard(w)
                    if not waiting:
                        launch(gi)
                        progressed = True
            if not progressed:
                # nothing ready: block briefly on the worker pipes of the advancing groups (or yield)
                fds = [pool[w] for gi in range(ngroups) if state[gi] == ADV for w in info[gi]]
                if fds:
                    select.select(fds, [], [], 2e-4)
    results = []
    for wk in pool:
        wk.send(("finish",))
    for wk in pool:
        X, F, G, TASK, NIT, NFEV = wk.recv()
        for b in range(len(F)):
            t0, t1 = int(TASK[b, 0]), int(TASK[b, 1])
            nit, nfev = int(NIT[b]), int(NFEV[b])
            warnflag = 0 if t0 == 4 else (1 if (nfev > maxfun or nit >= maxiter) else 2)
            msg = _lb.status_messages[t0] + ": " + _lb.task_messages[t1]
            results.append(scipy.optimize.OptimizeResult(fun=float(F[b]), jac=G[b].copy(), nfev=nfev, njev=nfev, nit=nit,
                                                         status=warnflag, message=msg, x=X[b].copy(), success=(warnflag == 0)))
    return results


# ---- multi-GPU partitioning -----------------------------------------------------------------------------


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition of ``total`` independent GPs: rank r owns [lo, hi)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather the per-rank result rows ([B_local, C]) into [total, C] (the single collective of
    the C3 path).  Works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


# ---- the batched model ------------------------------------------------------------------------------------


class BatchedGPR:
    """B independent exact GPs sharing one kernel expression.

    X [B,N,D], Y [B,N] (or [B,N,1]); ``kernel`` gives the expression, the trainable flags and the
    initial hyper-parameters of every GP; ``noise_variance`` scalar or [B] (the restart grid of
    models/model_trainer.py:26 is a [B] vector); ``train_noise`` mirrors
    ``set_trainable(model.likelihood, True/False)``.

    ``nrows`` [B] (optional): GP b is fitted on the first ``nrows[b]`` rows of ``X[b]`` only.  That is how
    the reference's rolling re-fit maps onto one batch WITHOUT changing the model: its loop
    (Multi-Input_GPR/main.py:414-456) fits EXPANDING windows ``X_full[:i]``, one more row per test day, so
    ``X[b] = X_full[:Nmax]`` for every b and ``nrows = [i0, i0 + 1, ...]`` (``data_prep.expanding_windows``
    builds exactly that) gives each GP the data the reference gives it.  ``data_prep.rolling_windows`` cuts
    fixed-length SLIDING windows instead (BASELINE config C3): a different model from the reference's loop.
    Windows of up to 128 rows run one GP per CTA with K in shared memory (a 128 x 132 fp64 tile is 135 of the
    227 KB).  LONGER windows (the reference's loop grows by one row per test day) take the blocked path of
    ``models.GPR`` instead, up to ``MANY_HANDLES`` (16) evaluations side by side: one engine handle, CUDA stream and
    host thread of the library each (``gpb_gpr_lml_grad_many``) -- a single evaluation at a few hundred rows is
    bound by the one-CTA chain of its factorisation and leaves the GPU idle (N = 256: 7.0 K evaluations/s one at a
    time, 35 K with 16 side by side; N = 1000: 2.5 K -> 7.5 K; profiles/r02_long_windows.json).  Same class, same methods, same
    lock-step L-BFGS-B; ``nrows`` makes the batch ragged on either path."""

    MANY_HANDLES = 16

    def __init__(self, X, Y, kernel: Kernel, noise_variance=1.0, train_noise: bool = True, device=None, nrows=None):
        self.device_index = ops.cuda_device_index(device)
        self.X = ops.to_device(X, self.device_index)
        if self.X.ndim != 3:
            raise ValueError("X must be [B, N, D]")
        self.B, self.N, self.D = (int(v) for v in self.X.shape)
        Yd = ops.to_device(Y, self.device_index)
        if Yd.ndim == 3:
            if Yd.shape[2] != 1:
                raise NotImplementedError("single-output GPs only")
            Yd = Yd[:, :, 0]
        if tuple(Yd.shape) != (self.B, self.N):
            raise ValueError("Y must be [B, N]")
        self.Y = Yd.contiguous()
        self._large = self.N > 128
        self.nrows = None
        self._nrows_dev = None
        if nrows is not None:
            nr = np.asarray(nrows, dtype=np.int32).reshape(-1)
            if nr.shape[0] != self.B or nr.min() < 1 or nr.max() > self.N:
                raise ValueError("nrows must be [B] with 1 <= nrows[b] <= N")
            self.nrows = nr
            self._nrows_dev = torch.from_numpy(nr).to(self.X.device)
        self.kernel = kernel
        self.compiled = compile_kernel(kernel, self.D)
        self.P = self.compiled.n_params
        theta0 = self.compiled.theta()
        self.theta = np.tile(theta0[None, :], (self.B, 1))                     # constrained, [B,P]
        self.noise = np.broadcast_to(np.asarray(noise_variance, dtype=np.float64), (self.B,)).copy()
        self.train_noise = bool(train_noise)
        # which theta slots are trainable (by the template kernel's Parameter flags)
        mask = np.zeros(self.P, dtype=bool)
        for p, o in zip(self.compiled.params, self.compiled.offsets):
            mask[o:o + max(1, p.size)] = p.trainable
        self.trainable_mask = mask
        self._theta_tf = Softplus(0.0)
        self._noise_tf = Softplus(DEFAULT_VARIANCE_LOWER_BOUND)
        self.non_pd_evaluations = np.zeros(self.B, dtype=np.int64)
        self._engine = ops.shared_engine(self.device_index)
        self._out = torch.empty((self.B, 2 + self.P), dtype=torch.float64, device=self.X.device)
        self._info = torch.zeros((self.B,), dtype=torch.int32, device=self.X.device)
        # further result sets for the pipelined fit (several part-batches in flight, lockstep_lbfgsb)
        self._more_out = {}
        self._slot_busy = [False]
        self._executor = None
        if self._large:
            # windows longer than 128 rows: engines of their own (workspaces, streams) for the side-by-side path
            ranks_here = 1
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    ranks_here = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", dist.get_world_size())))
            except Exception:
                pass
            cores = max(1, (os.cpu_count() or 1) // ranks_here)      # one library thread per handle
            nh = max(1, min(self.B, int(os.environ.get("GPB_MANY_HANDLES", min(self.MANY_HANDLES, cores)))))
            # every handle holds its own O(N^2) workspaces (about six N x N fp64 matrices): stay within a quarter
            # of the free device memory
            free_bytes, _ = torch.cuda.mem_get_info(self.X.device)
            nh = max(1, min(nh, int(0.25 * free_bytes / (6.0 * 8.0 * self.N * self.N))))
            self._engines = [_capi.Engine(self.device_index) for _ in range(nh)]
            self._streams = [torch.cuda.Stream(device=self.X.device) for _ in range(nh)]
            for e, st in zip(self._engines, self._streams):
                e.set_stream(st.cuda_stream)
                e.set_kernel(self.compiled.spec, self.compiled.token)
            esz = self.X.element_size()
            self._xptr = np.array([self.X.data_ptr() + b * self.N * self.D * esz for b in range(self.B)], dtype=np.uint64)
            self._yptr = np.array([self.Y.data_ptr() + b * self.N * esz for b in range(self.B)], dtype=np.uint64)
            self._nr64 = (self.nrows.astype(np.int64) if self.nrows is not None else np.full(self.B, self.N, dtype=np.int64))
            torch.cuda.synchronize(self.X.device)      # X, Y are complete before the engines' own streams read them

    def _many(self, theta: np.ndarray, noise: np.ndarray, idx: Optional[np.ndarray], want_grad: bool):
        """(lml, dlml/dtheta, dlml/dnoise, info) of the selected GPs through ``gpb_gpr_lml_grad_many``."""
        sel = np.arange(self.B) if idx is None else np.asarray(idx, dtype=np.int64)
        lml, g, gn, rc = _capi.Engine.gpr_lml_grad_many(self._engines, self._xptr[sel], self._nr64[sel], self.D, self._yptr[sel],
                                                        theta, noise, want_grad)
        return lml, g, gn, rc

    # -- raw device evaluation ---------------------------------------------------------------------
    def _launch(self, theta: np.ndarray, noise: np.ndarray, idx: Optional[np.ndarray], want_grad: bool, slot: int = 0):
        eng = self._engine
        ops.sync_stream(eng)
        eng.set_kernel(self.compiled.spec, self.compiled.token)
        dev = self.X.device
        th = torch.from_numpy(np.ascontiguousarray(theta, dtype=np.float64)).to(dev, non_blocking=True)
        nz = torch.from_numpy(np.ascontiguousarray(noise, dtype=np.float64)).to(dev, non_blocking=True)
        nr = self._nrows_dev
        if idx is None:
            Xb, Yb, b = self.X, self.Y, self.B
        else:
            it = torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int64)).to(dev)
            Xb, Yb, b = self.X.index_select(0, it).contiguous(), self.Y.index_select(0, it).contiguous(), len(idx)
            if nr is not None:
                nr = nr.index_select(0, it).contiguous()
        if slot == 0:
            out, info = self._out[:b], self._info[:b]
        else:
            if slot not in self._more_out:
                self._more_out[slot] = (torch.empty_like(self._out), torch.zeros_like(self._info))
            out, info = self._more_out[slot][0][:b], self._more_out[slot][1][:b]
        eng.batched_lml_grad(Xb.data_ptr(), Yb.data_ptr(), th.data_ptr(), nz.data_ptr(), b, self.N, self.D,
                             out.data_ptr(), info.data_ptr(), want_grad, None if nr is None else nr.data_ptr())
        # (the inputs th, nz, Xb, Yb stay referenced by the caller-visible tensors of this stream-ordered launch;
        # torch's caching allocator does not reuse them before the kernel has run on the same stream)
        return out, info

    def lml_and_grads(self, theta: Optional[np.ndarray] = None, noise: Optional[np.ndarray] = None,
                      idx: Optional[np.ndarray] = None, want_grad: bool = True, subset_params: bool = False):
        """(lml [b], dlml/dtheta [b,P] constrained, dlml/dnoise [b], info [b]) as numpy arrays.

        ``idx`` selects the GPs (rows of X, Y).  ``theta`` / ``noise`` are FULL-batch arrays ([B, P], [B]) that
        are indexed with ``idx`` here, unless ``subset_params=True``: then they already hold one row per entry
        of ``idx``, in that order (the lock-step driver's case).  Nothing is inferred from shapes."""
        theta = self.theta if theta is None else np.asarray(theta)
        noise = self.noise if noise is None else np.asarray(noise)
        if idx is not None:
            if subset_params:
                if theta.shape[0] != len(idx) or noise.shape[0] != len(idx):
                    raise ValueError("subset_params=True: theta and noise need one row per entry of idx")
            else:
                if theta.shape[0] != self.B or noise.shape[0] != self.B:
                    raise ValueError("theta and noise must be full-batch arrays ([B, P], [B]); pass subset_params=True for per-idx rows")
                theta, noise = theta[idx], noise[idx]
        if self._large:
            return self._many(theta, noise, idx, want_grad)
        out, info = self._launch(theta, noise, idx, want_grad)
        o = out.cpu().numpy()
        return o[:, 0].copy(), o[:, 2:].copy(), o[:, 1].copy(), info.cpu().numpy()

    # -- optimisation in unconstrained space ---------------------------------------------------------
    def _pack(self) -> np.ndarray:
        cols = [self._theta_tf.inverse(self.theta[:, self.trainable_mask])]
        if self.train_noise:
            cols.append(self._noise_tf.inverse(self.noise)[:, None])
        return np.concatenate(cols, axis=1)

    def _unpack(self, U: np.ndarray, idx: np.ndarray):
        nt = int(self.trainable_mask.sum())
        theta = self.theta[idx].copy()
        theta[:, self.trainable_mask] = self._theta_tf.forward(U[:, :nt])
        noise = self._noise_tf.forward(U[:, nt]) if self.train_noise else self.noise[idx].copy()
        return theta, noise

    # A trial point where K + s2 I is not positive definite: GPflow / TF raise out of ``minimize`` there
    # (InvalidArgumentError) and the whole fit is lost.  In a batch one bad GP must not take
    # the others down, so its objective is reported as this large FINITE value with a zero gradient: SciPy's
    # line search (dcsrch) backs off from it like from any too-long step, where +inf would turn its
    # interpolation into NaN.  The GP is recorded in ``self.non_pd_evaluations`` and its iterates may differ
    # from what a run that never hits such a point would give -- for those GPs the "iterates are SciPy's own"
    # guarantee is about the SciPy state machine, not about GPflow (which would have raised).
    NON_PD_PENALTY = 1e100

    def _finish_unconstrained(self, U, idx, lml, gth, gnz, info):
        nt = int(self.trainable_mask.sum())
        g = -gth[:, self.trainable_mask] * self._theta_tf.forward_grad(U[:, :nt])
        if self.train_noise:
            g = np.concatenate([g, (-gnz * self._noise_tf.forward_grad(U[:, nt]))[:, None]], axis=1)
        f = -lml
        bad = info != 0
        if np.any(bad):
            f = np.where(bad, self.NON_PD_PENALTY, f)
            g = np.where(bad[:, None], 0.0, g)
            np.add.at(self.non_pd_evaluations, np.asarray(idx)[bad], 1)
        return f, g

    def loss_and_grads_unconstrained(self, U: np.ndarray, idx: np.ndarray):
        """training_loss (= -LML) and its gradient w.r.t. the packed unconstrained variables."""
        theta, noise = self._unpack(U, idx)
        lml, gth, gnz, info = self.lml_and_grads(theta, noise, idx, subset_params=True)
        return self._finish_unconstrained(U, idx, lml, gth, gnz, info)

    def loss_and_grads_unconstrained_async(self, U: np.ndarray, idx: np.ndarray):
        """Start the evaluation (H2D copies and the kernel go onto the stream) and return ``wait() -> (f, g)``
        (``wait.ready()`` tells whether it would block); several evaluations may be in flight, each with its
        own result buffers."""
        U = np.array(U, dtype=np.float64)
        idx = np.array(idx, dtype=np.int64)
        if self._large:
            # the side-by-side path is a blocking library call (ctypes drops the GIL for its duration): it runs on
            # one helper thread, so that the caller can advance the SciPy state machines of the other part-batch
            # meanwhile; one call at a time -- the handles are not shared between calls
            if self._executor is None:
                from concurrent.futures import ThreadPoolExecutor
                self._executor = ThreadPoolExecutor(max_workers=1, thread_name_prefix="gpb-many")
            fut = self._executor.submit(self.loss_and_grads_unconstrained, U, idx)

            def done():
                return fut.result()

            done.ready = fut.done
            return done
        theta, noise = self._unpack(U, idx)
        slot = next((s for s, busy in enumerate(self._slot_busy) if not busy), None)
        if slot is None:
            slot = len(self._slot_busy)
            self._slot_busy.append(False)
        self._slot_busy[slot] = True
        out, info = self._launch(theta, noise, idx, True, slot=slot)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.X.device))

        def wait():
            ev.synchronize()
            o = out.cpu().numpy()
            inf = info.cpu().numpy()
            self._slot_busy[slot] = False
            return self._finish_unconstrained(U, idx, o[:, 0].copy(), o[:, 2:].copy(), o[:, 1].copy(), inf)

        wait.ready = ev.query
        return wait

    @staticmethod
    def default_workers(B: int) -> int:
        """Worker processes for the host side of a lock-step fit of B GPs: SciPy's ``setulb`` (GIL-bound, ~3 us
        per problem and round) is most of a large fit, so batches of >= 512 GPs spread it over the host cores
        this rank may use (cores // ranks on the node - 1, at most 12); smaller batches stay in process
        (measured on one B200, 5120 GPs: 3620 fits/s in process, 12 700 with 12 workers in four groups)."""
        if B < 512:
            return 0
        import os
        world = 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", dist.get_world_size())))
        except Exception:
            pass
        n = min(12, (os.cpu_count() or 1) // world - 1)
        return n if n >= 2 else 0

    def fit(self, maxiter: int = 15000, **lbfgs_kwargs) -> List[scipy.optimize.OptimizeResult]:
        """Scipy().minimize(model.training_loss, model.trainable_variables) for every GP, lock step.
        ``workers=k`` (k > 1) advances the SciPy state machines in k worker processes (same iterates, bit for
        bit); the default (``workers="auto"``) is ``default_workers(B)``."""
        if lbfgs_kwargs.get("workers", "auto") == "auto":
            lbfgs_kwargs["workers"] = self.default_workers(self.B)
        U0 = self._pack()
        self.non_pd_evaluations = np.zeros(self.B, dtype=np.int64)
        # (long windows: two part-batches in flight measured slower -- 0.088 s against 0.072 s for 32 windows of
        # 200 - 231 rows: each library call then has half the evaluations to spread over its handles; opt-in only)
        pipelined = lbfgs_kwargs.pop("pipelined", self.B >= 64 and not self._large)
        res = lockstep_lbfgsb(self.loss_and_grads_unconstrained, U0, maxiter=maxiter,
                              fun_batch_async=self.loss_and_grads_unconstrained_async if pipelined else None, **lbfgs_kwargs)
        U = np.stack([r.x for r in res])
        self.theta, self.noise = self._unpack(U, np.arange(self.B))
        return res

    def predict_f(self, Xnew):
        """Per-GP predict_f(Xnew[b], full_cov=False): Xnew [B,Ns,D] -> (mean [B,Ns], var [B,Ns]) on device."""
        Xs = ops.to_device(Xnew, self.device_index)
        if Xs.ndim != 3 or Xs.shape[0] != self.B or Xs.shape[2] != self.D:
            raise ValueError("Xnew must be [B, Ns, D]")
        Ns = int(Xs.shape[1])
        dev = self.X.device
        if self._large:
            Xs = Xs.contiguous()
            mean = torch.empty((self.B, Ns), dtype=torch.float64, device=dev)
            var = torch.empty((self.B, Ns), dtype=torch.float64, device=dev)
            torch.cuda.current_stream(dev).synchronize()         # Xs, mean, var exist before the engines' streams use them
            esz = Xs.element_size()
            xs_ptr = np.array([Xs.data_ptr() + b * Ns * self.D * esz for b in range(self.B)], dtype=np.uint64)
            m_ptr = np.array([mean.data_ptr() + b * Ns * esz for b in range(self.B)], dtype=np.uint64)
            v_ptr = np.array([var.data_ptr() + b * Ns * esz for b in range(self.B)], dtype=np.uint64)
            rc = _capi.Engine.gpr_predict_f_many(self._engines, self._xptr, self._nr64, self.D, self._yptr, self.theta, self.noise,
                                                 xs_ptr, Ns, m_ptr, v_ptr)
            bad = np.nonzero(rc > 0)[0]
            if bad.size:
                # a GP whose covariance is not positive definite does not take the batch down (as on the
                # one-GP-per-CTA path): its predictions are NaN
                it = torch.from_numpy(bad.astype(np.int64)).to(dev)
                mean[it] = float("nan")
                var[it] = float("nan")
            return mean, var
        eng = self._engine
        ops.sync_stream(eng)
        eng.set_kernel(self.compiled.spec, self.compiled.token)
        th = torch.from_numpy(np.ascontiguousarray(self.theta)).to(dev)
        nz = torch.from_numpy(np.ascontiguousarray(self.noise)).to(dev)
        mean = torch.empty((self.B, Ns), dtype=torch.float64, device=dev)
        var = torch.empty((self.B, Ns), dtype=torch.float64, device=dev)
        eng.batched_predict_f(self.X.data_ptr(), self.Y.data_ptr(), th.data_ptr(), nz.data_ptr(), self.B, self.N, self.D,
                              Xs.data_ptr(), Ns, mean.data_ptr(), var.data_ptr(), self._info.data_ptr(),
                              None if self._nrows_dev is None else self._nrows_dev.data_ptr())
        return mean, var
