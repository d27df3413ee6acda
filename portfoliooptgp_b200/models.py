"""gpflow.models.GPR / SVGP as the reference drives them.

Every numerical step runs in libgpb200 (sm_100a CUDA) through the C-ABI; this layer only keeps
the GPflow object model: Parameters, ``training_loss`` / ``maximum_log_likelihood_objective`` /
``log_marginal_likelihood`` / ``elbo``, ``predict_f`` / ``predict_y``, ``trainable_variables``.
Because there is no autodiff tape here, objectives are differentiated analytically on the device
and ``optimizers.Scipy`` asks the model for ``(loss, gradients)`` directly.

Reference call sites: GPR/model_trainer.py:15-20, GPR/predictor.py:6-7,
Multi-Input_GPR/main.py:421-434, Multi-Input_GPR/models/model_trainer.py:19-40,
test_scripts/SVGP.py:515-540.
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _capi, ops
from .base import Module, Parameter, Variable, triangular
from .config import default_jitter
from .kernels import CompiledKernel, Kernel, compile_kernel, structure_token
from .likelihoods import Gaussian
from .mean_functions import MeanFunction, Zero

OUTPUT_DEVICE = "cpu"  # predict_* return torch CPU fp64 tensors by default; "cuda" keeps them on device


def set_output_device(device: str) -> None:
    global OUTPUT_DEVICE
    assert device in ("cpu", "cuda")
    OUTPUT_DEVICE = device


def _out(t: torch.Tensor) -> torch.Tensor:
    return t if OUTPUT_DEVICE == "cuda" else t.cpu()


class LossClosure:
    """What ``training_loss_closure`` returns: callable (-> loss) that also knows how to produce
    gradients w.r.t. a list of unconstrained variables (the role of tf.GradientTape in GPflow)."""

    def __init__(self, model: "GPModel", data=None):
        self.model = model
        self.data = data

    def __call__(self):
        return self.model._training_loss(self.data)

    def value_and_grads(self, variables: Sequence[Variable]):
        return self.model._training_loss_and_grads(variables, self.data)


class GPModel(Module):
    def __init__(self, kernel: Kernel, likelihood, mean_function: Optional[MeanFunction] = None, device=None):
        if not isinstance(kernel, Kernel):
            raise TypeError("kernel must be a portfoliooptgp_b200.kernels.Kernel")
        self.kernel = kernel
        self.likelihood = likelihood
        self.mean_function = mean_function if mean_function is not None else Zero()
        self._device_index = ops.cuda_device_index(device)
        self._engine: Optional[_capi.Engine] = None
        self._compiled: Optional[CompiledKernel] = None
        self._compiled_token = None

    # engine / kernel lowering ---------------------------------------------------------------
    def _get_engine(self) -> _capi.Engine:
        # One engine (handle + grow-only workspaces) per device, shared by every model: the
        # reference builds a fresh GPR per kernel candidate / rolling window / restart
        # (GPR/model_trainer.py:15, Multi-Input_GPR/main.py:421), and each model re-binds its data and
        # kernel expression at every evaluation, so nothing model-specific lives in the handle.
        if self._engine is None:
            self._engine = ops.shared_engine(self._device_index)
        ops.sync_stream(self._engine)
        return self._engine

    def _lower_kernel(self, D: int) -> CompiledKernel:
        token = (structure_token(self.kernel, D),)
        if self._compiled is None or self._compiled_token != token:
            self._compiled = compile_kernel(self.kernel, D)
            self._compiled_token = token
        self._get_engine().set_kernel(self._compiled.spec, token)
        return self._compiled

    def __deepcopy__(self, memo):
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in vars(self).items():
            if k in ("_engine", "_compiled", "_compiled_token"):
                setattr(new, k, None)
            elif isinstance(v, torch.Tensor):
                setattr(new, k, v)  # device data is immutable here: share it
            elif k == "data":
                setattr(new, k, v)
            else:
                setattr(new, k, copy.deepcopy(v, memo))
        return new

    # GPflow objective surface -----------------------------------------------------------------
    def training_loss(self):
        """-(log marginal likelihood + log prior); no priors on this path."""
        return self._training_loss(None)

    def training_loss_closure(self, data=None, *, compile: bool = True) -> LossClosure:
        return LossClosure(self, data)

    def _training_loss(self, data):
        return -self._mll(data)

    def _mll(self, data):
        raise NotImplementedError

    def _training_loss_and_grads(self, variables: Sequence[Variable], data=None):
        raise NotImplementedError

    def predict_f(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):
        raise NotImplementedError

    def predict_y(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):
        """gpflow GPModel.predict_y: likelihood.predict_mean_and_var(predict_f(Xnew))
        (GPflow recomputes predict_f here, GPR/predictor.py:7 -- so does this)."""
        if full_cov or full_output_cov:
            raise NotImplementedError("The predict_y method currently supports only the argument values full_cov=False and full_output_cov=False")
        f_mean, f_var = self.predict_f(Xnew, full_cov=False)
        return self.likelihood.predict_mean_and_var(Xnew, f_mean, f_var)

    def predict_f_samples(self, *a, **k):
        raise NotImplementedError("predict_f_samples is not used on the reference path")

    def _grads_for(self, variables: Sequence[Variable], by_param: Dict[int, np.ndarray], sign: float):
        """Map {id(Parameter): d objective/d unconstrained} onto the requested variable list."""
        var_owner = {id(p.unconstrained_variable): p for p in self.parameters}
        out = []
        for v in variables:
            p = var_owner.get(id(v))
            if p is None or id(p) not in by_param:
                raise ValueError("a variable passed to the optimiser does not influence this model's objective "
                                 "(GPflow: 'gradients are None'); pass model.trainable_variables")
            out.append(sign * np.asarray(by_param[id(p)], dtype=np.float64).reshape(v.shape))
        return out


class GPR(GPModel):
    """Exact GP regression with a Gaussian likelihood (gpflow/models/gpr.py)."""

    def __init__(self, data, kernel: Kernel, mean_function: Optional[MeanFunction] = None,
                 noise_variance: Optional[float] = None, likelihood: Optional[Gaussian] = None, device=None):
        if likelihood is not None and noise_variance is not None:
            raise ValueError("only one of noise_variance and likelihood may be given")
        if likelihood is None:
            likelihood = Gaussian(1.0 if noise_variance is None else noise_variance)
        super().__init__(kernel, likelihood, mean_function, device)
        X, Y = data
        Xd = ops.to_device(X, self._device_index, ndim=2)
        Yd = ops.to_device(Y, self._device_index, ndim=2)
        if Xd.shape[0] != Yd.shape[0]:
            raise ValueError(f"X has {Xd.shape[0]} rows, Y has {Yd.shape[0]}")
        if Yd.shape[1] != 1:
            raise NotImplementedError("only single-output Y [N,1] is supported (R = 1 on every reference call site)")
        self.data = (Xd, Yd)
        self._Yc: Optional[torch.Tensor] = None

    # -- binding -------------------------------------------------------------------------------
    def _bind(self) -> Tuple[_capi.Engine, CompiledKernel]:
        Xd, Yd = self.data
        eng = self._get_engine()
        ck = self._lower_kernel(Xd.shape[1])
        if isinstance(self.mean_function, Zero):
            if self._Yc is None:
                self._Yc = Yd[:, 0].contiguous()
        else:
            self._Yc = (Yd - self.mean_function(Xd))[:, 0].contiguous()
        eng.gpr_set_data(Xd.data_ptr(), Xd.shape[0], Xd.shape[1], self._Yc.data_ptr())
        return eng, ck

    def _noise(self) -> float:
        return float(self.likelihood.variance.numpy())

    # -- objective -----------------------------------------------------------------------------
    def log_marginal_likelihood(self):
        eng, ck = self._bind()
        return torch.tensor(eng.gpr_lml(ck.theta(), self._noise()), dtype=torch.float64)

    def maximum_log_likelihood_objective(self):
        return self.log_marginal_likelihood()

    def _mll(self, data):
        if data is not None:
            raise ValueError("GPR holds its data internally; training_loss takes no data")
        return self.log_marginal_likelihood()

    def lml_and_constrained_grads(self):
        """(lml, d lml/d theta [constrained, engine order], d lml/d noise_variance)."""
        eng, ck = self._bind()
        return eng.gpr_lml_grad(ck.theta(), self._noise())

    def _training_loss_and_grads(self, variables: Sequence[Variable], data=None):
        eng, ck = self._bind()
        lml, g_theta, g_noise = eng.gpr_lml_grad(ck.theta(), self._noise())
        by_param = ck.scatter_grad(g_theta)
        pv = self.likelihood.variance
        by_param[id(pv)] = np.asarray(g_noise) * pv.transform.forward_grad(pv.unconstrained_variable._value)
        if any(p.trainable for p in self.mean_function.parameters):
            raise NotImplementedError("trainable mean-function parameters are not supported yet")
        return -lml, self._grads_for(variables, by_param, -1.0)

    # -- prediction ----------------------------------------------------------------------------
    def predict_f(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):
        if full_cov or full_output_cov:
            raise NotImplementedError("predict_f(full_cov=True) is not on the reference path (GPR/predictor.py:6 uses full_cov=False)")
        eng, ck = self._bind()
        Xs = ops.to_device(Xnew, self._device_index, ndim=2)
        if Xs.shape[1] != self.data[0].shape[1]:
            raise ValueError(f"Xnew has {Xs.shape[1]} columns, the model was built with {self.data[0].shape[1]}")
        Ns = Xs.shape[0]
        out = torch.empty((2, Ns), dtype=torch.float64, device=Xs.device)
        eng.gpr_predict_f(ck.theta(), self._noise(), Xs.data_ptr(), Ns, out[0].data_ptr(), out[1].data_ptr())
        mean = out[0][:, None]
        if not isinstance(self.mean_function, Zero):
            mean = mean + self.mean_function(Xs)
        return _out(mean), _out(out[1][:, None])
