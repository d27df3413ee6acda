"""The reference's two ``ModelTrainer`` classes with their INDEPENDENT
fits in flight at the same time on one GPU.

  ``GPRModelTrainer``        = GPR/model_trainer.py:6-26           (one fit per kernel candidate, MSE select)
  ``MultiInputModelTrainer`` = Multi-Input_GPR/models/model_trainer.py:8-72
                               (``train_model``, ``train_likelihood`` over the restart grid, ``train_best_model``)

Same method names, argument order, return values and per-model arithmetic: every fit is the reference's own
``Scipy().minimize(model.training_loss, model.trainable_variables, ...)`` on its own model, so its iterates are
exactly those of the sequential loop.  What changes is the schedule: each fit runs on its own host thread with
its own engine handle and CUDA stream (``fit_concurrently``).  A single exact-GP evaluation at these sizes is
bound by the one-CTA pivot chain of the blocked Cholesky, not by the machine: with four
C1-size fits in flight the GPU completes 6686 LML+gradient evaluations per second instead of 2122 one after
the other, at N = 8192 52.5 instead of 48.4 (tools/concurrent_evals.py, profiles/r02_concurrent_evals.json).
The reference's loops are sequential; nothing in them depends on the order (each candidate / restart gets its
own kernel object -- candidates that SHARE a kernel instance are fitted one after the other, in list order,
as the reference does).
"""
from __future__ import annotations

import threading
from copy import deepcopy
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from . import _capi, ops
from .models import GPR
from .optimizers import Scipy
from .utilities import print_summary, set_trainable

_pool_lock = threading.Lock()
_pool = {}   # device index -> list of idle (engine, stream) pairs


def _acquire(device_index: int):
    with _pool_lock:
        idle = _pool.setdefault(device_index, [])
        if idle:
            return idle.pop()
    return _capi.Engine(device_index), torch.cuda.Stream(device=device_index)


def _release(device_index: int, pair) -> None:
    with _pool_lock:
        _pool.setdefault(device_index, []).append(pair)


def run_concurrently(tasks: Sequence[Callable[[], object]], device_index: int, models_of_task=None, max_workers: int = 4) -> List[object]:
    """Run independent callables, at most ``max_workers`` at a time, each on its own host thread, engine handle
    and CUDA stream.  ``models_of_task[i]`` lists the models task i touches (their engine is switched to the
    thread's handle for the duration).  Exceptions are re-raised in the caller, first task first."""
    n = len(tasks)
    results: List[object] = [None] * n
    errors: List[Optional[BaseException]] = [None] * n
    torch.cuda.synchronize(device_index)          # inputs built on the caller's stream are complete
    next_task = [0]
    lock = threading.Lock()

    def worker():
        pair = _acquire(device_index)
        eng, stream = pair
        try:
            with torch.cuda.device(device_index), torch.cuda.stream(stream):
                while True:
                    with lock:
                        i = next_task[0]
                        next_task[0] += 1
                    if i >= n:
                        break
                    touched = list(models_of_task[i]) if models_of_task is not None else []
                    saved = [(m, m._engine, getattr(m, "_fact", None)) for m in touched]
                    try:
                        for m in touched:
                            m._engine = eng
                            if hasattr(m, "_fact"):
                                m._fact = None
                        results[i] = tasks[i]()
                    except BaseException as e:   # noqa: BLE001 - re-raised in the caller
                        errors[i] = e
                    finally:
                        stream.synchronize()
                        for m, e0, _ in saved:
                            m._engine = e0       # back on the shared handle of its device
                            if hasattr(m, "_fact"):
                                m._fact = None   # the factorisation stayed in the worker's handle
        finally:
            _release(device_index, pair)

    threads = [threading.Thread(target=worker, daemon=True) for _ in range(max(1, min(max_workers, n)))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None:
            raise e
    return results


def fit_concurrently(models: Sequence, options: Optional[dict] = None, max_workers: int = 4, after_fit: Optional[Callable] = None):
    """``Scipy().minimize(m.training_loss, m.trainable_variables, options=options)`` for every model, the
    independent ones at the same time.  Models that share a parameter object (e.g. the same kernel instance)
    form one group and are fitted in list order inside it, exactly as a sequential loop would.
    ``after_fit(model)`` (optional) runs right after a model's fit on the same thread (in-sample prediction, ...).
    Returns ``[(OptimizeResult, after_fit result)]`` in the order of ``models``."""
    models = list(models)
    if not models:
        return []
    # union-find over shared parameter objects
    owner = {}
    parent = list(range(len(models)))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for i, m in enumerate(models):
        for p in m.parameters:
            j = owner.setdefault(id(p), i)
            if j != i:
                parent[find(i)] = find(j)
    groups = {}
    for i in range(len(models)):
        groups.setdefault(find(i), []).append(i)
    order = sorted(groups.values(), key=lambda g: g[0])
    out = [None] * len(models)

    def make_task(idxs):
        def task():
            for i in idxs:
                m = models[i]
                res = Scipy().minimize(m.training_loss, m.trainable_variables, options=dict(options or {}))
                extra = after_fit(m) if after_fit is not None else None
                out[i] = (res, extra)
            return None
        return task

    dev = models[0]._device_index
    run_concurrently([make_task(g) for g in order], dev, models_of_task=[[models[i] for i in g] for g in order],
                     max_workers=max_workers)
    return out


def _mse(Y, mean) -> float:
    # sklearn.metrics.mean_squared_error on [N, 1] columns
    y = Y.detach().cpu().numpy() if hasattr(Y, "detach") else np.asarray(Y, dtype=np.float64)
    m = mean.detach().cpu().numpy() if hasattr(mean, "detach") else np.asarray(mean, dtype=np.float64)
    return float(np.mean((y.reshape(-1) - m.reshape(-1)) ** 2))


class GPRModelTrainer:
    """GPR/model_trainer.py:6-26.  ``max_workers`` fits in flight (1 = the reference's sequential loop)."""

    def __init__(self, kernel_combinations, max_workers: int = 4):
        self.kernel_combinations = kernel_combinations
        self.max_workers = max_workers

    def train_model(self, X_tf, Y_tf):
        models = []
        for kernel in self.kernel_combinations:
            model = GPR(data=(X_tf, Y_tf), kernel=kernel)                 # :15
            model.likelihood.variance.assign(1e-5)                       # :16
            set_trainable(model.likelihood.variance, False)              # :17
            models.append(model)
        fitted = fit_concurrently(models, options=dict(maxiter=100), max_workers=self.max_workers,    # :18-19
                                  after_fit=lambda m: _mse(Y_tf, m.predict_f(X_tf)[0]))                # :20-21
        best_kernel, best_mse, best_model = None, float("inf"), None
        for kernel, model, (_, mse_test) in zip(self.kernel_combinations, models, fitted):            # :22-25, list order
            if mse_test < best_mse:
                best_mse, best_kernel, best_model = mse_test, kernel, model
        return best_kernel, best_mse, best_model


class MultiInputModelTrainer:
    """Multi-Input_GPR/models/model_trainer.py:8-72.  ``train_model`` and ``train_likelihood`` are called on the
    class in the reference (no ``self``), so they are static here too."""

    max_workers = 4

    def __init__(self, kernel_combinations, max_workers: int = 4):
        self.kernel_combinations = kernel_combinations
        self.max_workers = max_workers

    @staticmethod
    def train_model(model):
        set_trainable(model.likelihood, False)                                                   # :19
        Scipy().minimize(model.training_loss, model.trainable_variables)                         # :20-21
        print_summary(model)                                                                     # :22
        return model

    @staticmethod
    def train_models(models: Sequence, max_workers: int = 4, summary: bool = False):
        """``train_model`` for several independent models at once (the per-asset / per-day fits of
        Multi-Input_GPR/main.py:414-456 when the windows are too long for ``BatchedGPR``)."""
        for m in models:
            set_trainable(m.likelihood, False)
        fit_concurrently(models, max_workers=max_workers)
        if summary:
            for m in models:
                print_summary(m)
        return list(models)

    @staticmethod
    def train_likelihood(X, Y, composite_kernel, starting_variances=(1e-5, 1e-3, 1e-1, 1.0), max_workers: int = 4, summary: bool = True):
        models = []
        for start_var in starting_variances:                                                     # :29-31
            model = GPR((X, Y), kernel=deepcopy(composite_kernel), noise_variance=start_var)
            set_trainable(model.likelihood, True)                                                # :34
            models.append(model)
        fitted = fit_concurrently(models, max_workers=max_workers)                               # :36-37, the restarts at once
        best_model, best_loss = None, float("inf")
        for model, (opt_logs, _) in zip(models, fitted):                                         # :40-48, grid order
            if opt_logs.fun < best_loss:
                best_model, best_loss = model, opt_logs.fun
        if summary:
            print("\nBest model:")
            print_summary(best_model)
            print(f"Best loss: {best_loss}")
        return best_model

    def train_best_model(self, X_tf, Y_tf):
        return GPRModelTrainer(self.kernel_combinations, self.max_workers).train_model(X_tf, Y_tf)   # :56-72 = R1
