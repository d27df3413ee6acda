"""gpflow.optimizers.Scipy: SciPy L-BFGS-B stays the host-side outer loop.

``minimize(closure, variables, **scipy_kwargs)`` packs the UNCONSTRAINED variables in the order
given, calls ``scipy.optimize.minimize(fun, x0, jac=True, method="L-BFGS-B", ...)`` (SciPy defaults
maxcor 10, ftol 2.22e-9, gtol 1e-5, maxiter = maxfun = 15000, maxls 20; the reference passes
``options=dict(maxiter=100)`` at GPR/model_trainer.py:19 and nothing at
Multi-Input_GPR/models/model_trainer.py:21), assigns ``result.x`` back and returns the
``OptimizeResult`` (``.fun`` is read at models/model_trainer.py:40).
Each objective evaluation is one device pass (loss + analytic gradient) and one stream sync.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import scipy.optimize

from .base import Variable
from .models import GPModel, LossClosure


def _as_closure(closure) -> LossClosure:
    if isinstance(closure, LossClosure):
        return closure
    owner = getattr(closure, "__self__", None)
    if isinstance(owner, GPModel) and getattr(closure, "__name__", "") == "training_loss":
        return LossClosure(owner, None)
    raise TypeError(
        "Scipy.minimize needs model.training_loss or model.training_loss_closure(data): this engine "
        "differentiates the model objective analytically on the GPU and cannot trace arbitrary Python closures")


class Scipy:
    def minimize(self, closure, variables: Sequence[Variable], method: Optional[str] = "L-BFGS-B",
                 step_callback: Optional[Callable] = None, compile: bool = True, allow_unused_variables: bool = False,
                 tf_fun_args=None, track_loss_history: bool = False, **scipy_kwargs) -> scipy.optimize.OptimizeResult:
        lc = _as_closure(closure)
        variables = tuple(variables)
        if not variables:
            raise ValueError("no variables to optimise")
        if not all(isinstance(v, Variable) for v in variables):
            raise TypeError("variables must be the model's (unconstrained) trainable_variables")
        x0 = self.initial_parameters(variables)
        history: List[float] = []

        def fun(x):
            self.assign_tensors(variables, x)
            loss, grads = lc.value_and_grads(variables)
            if track_loss_history:
                history.append(float(loss))
            return float(loss), self.pack_tensors(grads)

        callback = None
        if step_callback is not None:
            step = [0]

            def callback(x):
                self.assign_tensors(variables, x)
                step_callback(step[0], variables, [v.numpy() for v in variables])
                step[0] += 1

        if "callback" in scipy_kwargs and callback is not None:
            raise ValueError("Callback passed both via `step_callback` and `callback`")
        if callback is not None:
            scipy_kwargs["callback"] = callback
        result = scipy.optimize.minimize(fun, x0, jac=True, method=method, **scipy_kwargs)
        self.assign_tensors(variables, result.x)
        if track_loss_history:
            result["loss_history"] = history
        return result

    @staticmethod
    def initial_parameters(variables: Sequence[Variable]) -> np.ndarray:
        return np.concatenate([np.asarray(v._value, dtype=np.float64).reshape(-1) for v in variables])

    @staticmethod
    def pack_tensors(tensors) -> np.ndarray:
        return np.concatenate([np.asarray(t, dtype=np.float64).reshape(-1) for t in tensors])

    @staticmethod
    def assign_tensors(variables: Sequence[Variable], x: np.ndarray) -> None:
        pos = 0
        for v in variables:
            n = v.size
            v.assign(np.asarray(x[pos:pos + n], dtype=np.float64).reshape(v.shape))
            pos += n
        if pos != len(x):
            raise ValueError("packed vector length does not match the variables")
