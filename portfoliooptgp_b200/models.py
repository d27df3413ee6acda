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
            if k in ("_engine", "_compiled", "_compiled_token", "_fact"):
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

    # (SVGP overrides training_loss / _training_loss: its objective needs the minibatch)

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
        # (engine, factorisation serial) of this model's own last evaluation: predict_f hands it back so that
        # the engine can reuse W = L^-1 when theta / noise are unchanged (predict_y after predict_f, predict
        # after the last objective evaluation).  Safe because this object keeps X alive and immutable.
        self._fact: Optional[Tuple[int, int]] = None

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
    def _remember_factor(self, eng):
        self._fact = (id(eng), eng.gpr_factor_serial()) if hasattr(eng, "gpr_factor_serial") else None

    def log_marginal_likelihood(self):
        eng, ck = self._bind()
        lml = eng.gpr_lml(ck.theta(), self._noise())
        self._remember_factor(eng)
        return torch.tensor(lml, dtype=torch.float64)

    def maximum_log_likelihood_objective(self):
        return self.log_marginal_likelihood()

    def _mll(self, data):
        if data is not None:
            raise ValueError("GPR holds its data internally; training_loss takes no data")
        return self.log_marginal_likelihood()

    def lml_and_constrained_grads(self):
        """(lml, d lml/d theta [constrained, engine order], d lml/d noise_variance)."""
        eng, ck = self._bind()
        out = eng.gpr_lml_grad(ck.theta(), self._noise())
        self._remember_factor(eng)
        return out

    def _training_loss_and_grads(self, variables: Sequence[Variable], data=None):
        eng, ck = self._bind()
        lml, g_theta, g_noise = eng.gpr_lml_grad(ck.theta(), self._noise())
        self._remember_factor(eng)
        by_param = ck.scatter_grad(g_theta)
        pv = self.likelihood.variance
        by_param[id(pv)] = np.asarray(g_noise) * pv.transform.forward_grad(pv.unconstrained_variable._value)
        if any(p.trainable for p in self.mean_function.parameters):
            # dLML/dm(X) = alpha = (K + s2 I)^-1 (Y - m(X)); chain rule through the mean function on the device
            Xd = self.data[0]
            alpha = torch.empty(Xd.shape[0], dtype=torch.float64, device=Xd.device)
            eng.gpr_get_alpha(alpha.data_ptr())
            by_param.update(self.mean_function.backward(Xd, alpha))
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
        if self._fact is not None and self._fact[0] == id(eng) and hasattr(eng, "gpr_predict_f_reuse"):
            eng.gpr_predict_f_reuse(ck.theta(), self._noise(), self._fact[1], Xs.data_ptr(), Ns, out[0].data_ptr(),
                                    out[1].data_ptr())
        else:
            eng.gpr_predict_f(ck.theta(), self._noise(), Xs.data_ptr(), Ns, out[0].data_ptr(), out[1].data_ptr())
        self._remember_factor(eng)
        mean = out[0][:, None]
        if not isinstance(self.mean_function, Zero):
            mean = mean + self.mean_function(Xs)
        return _out(mean), _out(out[1][:, None])


class InducingPoints(Module):
    """gpflow.inducing_variables.InducingPoints: Z [M, D], trainable, identity transform."""

    def __init__(self, Z):
        Z = np.asarray(Z.numpy() if isinstance(Z, Parameter) else ops_to_numpy(Z), dtype=np.float64)
        if Z.ndim == 1:
            Z = Z[:, None]
        self.Z = Parameter(Z, name="Z")

    @property
    def num_inducing(self) -> int:
        return int(self.Z.shape[0])

    def __len__(self):
        return self.num_inducing


def ops_to_numpy(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


class SVGP(GPModel):
    """Sparse variational GP (gpflow/models/svgp.py) as constructed at test_scripts/SVGP.py:515-521:
    ``SVGP(kernel, Gaussian(variance), inducing_variable=Z, num_data=N)`` with GPflow's defaults
    whiten=True, q_diag=False, one latent GP, q_mu = 0, q_sqrt = I.

    elbo(data) = num_data / B * sum_b E_q[log p(y_b | f_b)] - KL[q(u) || p(u)]  (SURVEY.md G13-G14);
    the data term, its adjoint and the KL run on the device (csrc/svgp.cu)."""

    def __init__(self, kernel: Kernel, likelihood: Gaussian, inducing_variable, *, mean_function=None,
                 num_latent_gps: int = 1, q_diag: bool = False, q_mu=None, q_sqrt=None, whiten: bool = True,
                 num_data: Optional[int] = None, device=None):
        if not isinstance(likelihood, Gaussian):
            raise NotImplementedError("only the Gaussian likelihood is on the reference path")
        if num_latent_gps != 1 or q_diag or not whiten:
            raise NotImplementedError("SVGP supports the reference configuration: one latent GP, q_diag=False, whiten=True")
        if mean_function is not None and not isinstance(mean_function, Zero):
            raise NotImplementedError("SVGP mean functions are not on the reference path")
        super().__init__(kernel, likelihood, mean_function, device)
        self.inducing_variable = inducing_variable if isinstance(inducing_variable, InducingPoints) else InducingPoints(inducing_variable)
        M = self.inducing_variable.num_inducing
        self.num_data = num_data
        self.whiten = whiten
        self.q_diag = q_diag
        self.num_latent_gps = 1
        self.q_mu = Parameter(np.zeros((M, 1)) if q_mu is None else np.asarray(ops_to_numpy(q_mu), dtype=np.float64).reshape(M, 1),
                              name="q_mu")
        qs = np.eye(M)[None] if q_sqrt is None else np.asarray(ops_to_numpy(q_sqrt), dtype=np.float64).reshape(1, M, M)
        self.q_sqrt = Parameter(qs, transform=triangular(M), name="q_sqrt")
        self._flat: Optional[torch.Tensor] = None

    def _children(self):
        for key, val in super()._children():
            if key not in ("num_data", "whiten", "q_diag", "num_latent_gps"):
                yield key, val

    # -- helpers ---------------------------------------------------------------------------------
    def _device_params(self):
        dev = torch.device("cuda", self._device_index)
        Z = torch.from_numpy(np.ascontiguousarray(self.inducing_variable.Z.numpy())).to(dev)
        qmu = torch.from_numpy(np.ascontiguousarray(self.q_mu.numpy()[:, 0])).to(dev)
        Lq = torch.from_numpy(np.ascontiguousarray(np.tril(self.q_sqrt.numpy()[0]))).to(dev)
        return Z, qmu, Lq

    def _data(self, data):
        if data is None:
            raise ValueError("SVGP objectives need data=(X, Y)")
        X, Y = data
        Xd = ops.to_device(X, self._device_index, ndim=2)
        Yd = ops.to_device(Y, self._device_index, ndim=2)
        if Yd.shape[1] != 1 or Yd.shape[0] != Xd.shape[0]:
            raise ValueError("Y must be [N, 1] with as many rows as X")
        return Xd, Yd[:, 0].contiguous()

    def _scale(self, B: int) -> float:
        return 1.0 if self.num_data is None else float(self.num_data) / float(B)

    def _run(self, data, want_grad: bool):
        Xd, yd = self._data(data)
        eng = self._get_engine()
        Z, qmu, Lq = self._device_params()
        M, D = Z.shape
        if Xd.shape[1] != D:
            raise ValueError(f"X has {Xd.shape[1]} columns, Z has {D}")
        ck = self._lower_kernel(D)
        n = eng.svgp_flat_size(M, D, ck.n_params)
        if self._flat is None or self._flat.numel() != n:
            self._flat = torch.empty(n, dtype=torch.float64, device=Xd.device)
        noise = float(self.likelihood.variance.numpy())
        eng.svgp_data_term(ck.theta(), noise, Z.data_ptr(), M, D, qmu.data_ptr(), Lq.data_ptr(), M, Xd.data_ptr(),
                           yd.data_ptr(), Xd.shape[0], self._flat.data_ptr(), want_grad)
        elbo, kl = eng.svgp_finish(self._flat.data_ptr(), self._scale(Xd.shape[0]), qmu.data_ptr(), Lq.data_ptr(), M, M, D,
                                   ck.n_params, want_grad)
        return elbo, kl, ck, M, D

    # -- GPflow surface --------------------------------------------------------------------------
    def elbo(self, data):
        return torch.tensor(self._run(data, False)[0], dtype=torch.float64)

    def maximum_log_likelihood_objective(self, data):
        return self.elbo(data)

    def prior_kl(self):
        eng = self._get_engine()
        Z, qmu, Lq = self._device_params()
        M, D = Z.shape
        scratch = torch.zeros(eng.svgp_flat_size(M, D, 1), dtype=torch.float64, device=Z.device)
        return torch.tensor(eng.svgp_finish(scratch.data_ptr(), 1.0, qmu.data_ptr(), Lq.data_ptr(), M, M, D, 1, False)[1],
                            dtype=torch.float64)

    def training_loss(self, data):
        return -self.elbo(data)

    def training_loss_closure(self, data, *, compile: bool = True) -> LossClosure:
        Xd, yd = self._data(data)
        return LossClosure(self, (Xd, yd[:, None]))

    def _training_loss(self, data):
        return -self.elbo(data)

    def _training_loss_and_grads(self, variables: Sequence[Variable], data=None):
        elbo, kl, ck, M, D = self._run(data, True)
        flat = self._flat.cpu().numpy()
        P = ck.n_params
        by_param = ck.scatter_grad(flat[2:2 + P])
        pv = self.likelihood.variance
        by_param[id(pv)] = np.asarray(flat[1]) * pv.transform.forward_grad(pv.unconstrained_variable._value)
        o = 2 + P
        by_param[id(self.inducing_variable.Z)] = flat[o:o + M * D].reshape(M, D)
        o += M * D
        by_param[id(self.q_mu)] = flat[o:o + M].reshape(M, 1)
        o += M
        by_param[id(self.q_sqrt)] = self.q_sqrt.transform.pull_back(flat[o:o + M * M].reshape(1, M, M))
        return -elbo, self._grads_for(variables, by_param, -1.0)

    def predict_f(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):
        if full_cov or full_output_cov:
            raise NotImplementedError("predict_f(full_cov=True) is not on the reference path")
        eng = self._get_engine()
        Z, qmu, Lq = self._device_params()
        M, D = Z.shape
        Xs = ops.to_device(Xnew, self._device_index, ndim=2)
        if Xs.shape[1] != D:
            raise ValueError(f"Xnew has {Xs.shape[1]} columns, Z has {D}")
        ck = self._lower_kernel(D)
        Ns = Xs.shape[0]
        out = torch.empty((2, Ns), dtype=torch.float64, device=Xs.device)
        eng.svgp_predict_f(ck.theta(), Z.data_ptr(), M, D, qmu.data_ptr(), Lq.data_ptr(), M, Xs.data_ptr(), Ns,
                           out[0].data_ptr(), out[1].data_ptr())
        return _out(out[0][:, None]), _out(out[1][:, None])


class SGPR(GPModel):
    """Sparse GP regression with Titsias' collapsed bound (gpflow/models/sgpr.py), as constructed at
    test_scripts/SVGP.py:393-399: ``SGPR((X, Y), kernel, inducing_variable=Z)`` trained with Scipy on
    ``training_loss`` and queried with ``predict_y``.  The bound, its hand-derived adjoint w.r.t.
    kernel hyper-parameters, noise variance, inducing points and mean-function output, and predict_f
    run on the device (csrc/svgp.cu, gpb_sgpr_*)."""

    def __init__(self, data, kernel: Kernel, inducing_variable, *, mean_function: Optional[MeanFunction] = None,
                 num_latent_gps: Optional[int] = None, noise_variance: Optional[float] = None,
                 likelihood: Optional[Gaussian] = None, device=None):
        if likelihood is not None and noise_variance is not None:
            raise ValueError("only one of noise_variance and likelihood may be given")
        if likelihood is None:
            likelihood = Gaussian(1.0 if noise_variance is None else noise_variance)
        if num_latent_gps not in (None, 1):
            raise NotImplementedError("only single-output Y [N,1] is supported")
        super().__init__(kernel, likelihood, mean_function, device)
        X, Y = data
        Xd = ops.to_device(X, self._device_index, ndim=2)
        Yd = ops.to_device(Y, self._device_index, ndim=2)
        if Xd.shape[0] != Yd.shape[0]:
            raise ValueError(f"X has {Xd.shape[0]} rows, Y has {Yd.shape[0]}")
        if Yd.shape[1] != 1:
            raise NotImplementedError("only single-output Y [N,1] is supported (R = 1 on every reference call site)")
        self.data = (Xd, Yd)
        self.num_latent_gps = 1
        self.inducing_variable = inducing_variable if isinstance(inducing_variable, InducingPoints) else InducingPoints(inducing_variable)
        if self.inducing_variable.Z.shape[1] != Xd.shape[1]:
            raise ValueError(f"Z has {self.inducing_variable.Z.shape[1]} columns, X has {Xd.shape[1]}")

    def _children(self):
        for key, val in super()._children():
            if key != "num_latent_gps":
                yield key, val

    def _inputs(self):
        Xd, Yd = self.data
        eng = self._get_engine()
        ck = self._lower_kernel(Xd.shape[1])
        err = Yd[:, 0] if isinstance(self.mean_function, Zero) else (Yd - self.mean_function(Xd))[:, 0]
        Z = torch.from_numpy(np.ascontiguousarray(self.inducing_variable.Z.numpy())).to(Xd.device)
        return eng, ck, Xd, err.contiguous(), Z

    def _noise(self) -> float:
        return float(self.likelihood.variance.numpy())

    def elbo(self):
        eng, ck, Xd, err, Z = self._inputs()
        M, D = Z.shape
        out = eng.sgpr_elbo(ck.theta(), self._noise(), Z.data_ptr(), M, D, Xd.data_ptr(), err.data_ptr(), Xd.shape[0],
                            ck.n_params, False)
        return torch.tensor(out[0], dtype=torch.float64)

    def maximum_log_likelihood_objective(self):
        return self.elbo()

    def _mll(self, data):
        if data is not None:
            raise ValueError("SGPR holds its data internally; training_loss takes no data")
        return self.elbo()

    def _training_loss_and_grads(self, variables: Sequence[Variable], data=None):
        eng, ck, Xd, err, Z = self._inputs()
        M, D = Z.shape
        N = Xd.shape[0]
        train_mean = any(p.trainable for p in self.mean_function.parameters)
        errbar = torch.empty(N, dtype=torch.float64, device=Xd.device) if train_mean else None
        out = eng.sgpr_elbo(ck.theta(), self._noise(), Z.data_ptr(), M, D, Xd.data_ptr(), err.data_ptr(), N, ck.n_params,
                            True, None if errbar is None else errbar.data_ptr())
        P = ck.n_params
        by_param = ck.scatter_grad(out[2:2 + P])
        pv = self.likelihood.variance
        by_param[id(pv)] = np.asarray(out[1]) * pv.transform.forward_grad(pv.unconstrained_variable._value)
        by_param[id(self.inducing_variable.Z)] = out[2 + P:].reshape(M, D)
        if train_mean:
            # err = Y - m(X): d elbo/d m(X) = -err_bar
            by_param.update(self.mean_function.backward(Xd, -errbar))
        return -out[0], self._grads_for(variables, by_param, -1.0)

    def predict_f(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):
        if full_cov or full_output_cov:
            raise NotImplementedError("predict_f(full_cov=True) is not on the reference path")
        eng, ck, Xd, err, Z = self._inputs()
        M, D = Z.shape
        Xs = ops.to_device(Xnew, self._device_index, ndim=2)
        if Xs.shape[1] != D:
            raise ValueError(f"Xnew has {Xs.shape[1]} columns, the model was built with {D}")
        Ns = Xs.shape[0]
        out = torch.empty((2, Ns), dtype=torch.float64, device=Xs.device)
        eng.sgpr_predict_f(ck.theta(), self._noise(), Z.data_ptr(), M, D, Xd.data_ptr(), err.data_ptr(), Xd.shape[0],
                           Xs.data_ptr(), Ns, out[0].data_ptr(), out[1].data_ptr())
        mean = out[0][:, None]
        if not isinstance(self.mean_function, Zero):
            mean = mean + self.mean_function(Xs)
        return _out(mean), _out(out[1][:, None])
