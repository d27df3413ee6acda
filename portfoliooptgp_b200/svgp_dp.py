"""Data-parallel minibatch SVGP (north_star subsystem 5, BASELINE config C5).

One process per GPU.  Parameters are replicated; every rank owns a shard of the N rows
(pre-shuffled, so consecutive windows of the shard are uniform minibatches and are handed to the
engine zero-copy as pointer offsets), evaluates the UNSCALED data term and its gradient on its
minibatch, and the flat records are summed with ONE ``all_reduce`` (NCCL over NVLink; gloo in the
CPU tests).  Then, identically on every rank: scale by num_data / (world * B), subtract the KL term
and its gradient once, and step the optimiser (Adam, as GPflow users do for
minibatch SVGP; the variational parameters and Z are updated on the device, the few constrained
hyper-parameters on the host)."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import ops
from .base import Softplus
from .kernels import Kernel, compile_kernel
from .likelihoods import DEFAULT_VARIANCE_LOWER_BOUND


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def combine_records(S_sum: float, kl: float, num_data: float, world: int, B: int) -> float:
    """ELBO from the all-reduced data-term sum: scale once, subtract the KL once."""
    return float(num_data) / float(world * B) * S_sum - kl


class SVGPDataParallel:
    def __init__(self, kernel: Kernel, noise_variance: float, Z, num_data: int, X_shard, Y_shard, minibatch_size: int,
                 lr: float = 1e-2, train_hyper: bool = True, device=None, group=None):
        self.device_index = ops.cuda_device_index(device)
        self.dev = torch.device("cuda", self.device_index)
        self.group = group
        self.X = ops.to_device(X_shard, self.device_index, ndim=2)
        Y = ops.to_device(Y_shard, self.device_index, ndim=2)
        self.y = Y[:, 0].contiguous()
        self.D = int(self.X.shape[1])
        self.B = int(minibatch_size)
        if self.B > self.X.shape[0]:
            raise ValueError("minibatch larger than the local shard")
        self.num_data = int(num_data)
        self.kernel = kernel
        self.ck = compile_kernel(kernel, self.D)
        self.P = self.ck.n_params
        self.theta = self.ck.theta()
        self.noise = float(noise_variance)
        Zd = ops.to_device(Z, self.device_index, ndim=2)
        self.M = int(Zd.shape[0])
        self.engine = ops.shared_engine(self.device_index)
        n = self.engine.svgp_flat_size(self.M, self.D, self.P)
        self.flat = torch.zeros(n, dtype=torch.float64, device=self.dev)
        # device-resident parameters, laid out like the gradient slots of the flat record: Z, q_mu, q_sqrt
        self.nvar = self.M * self.D + self.M + self.M * self.M
        self.params = torch.zeros(self.nvar, dtype=torch.float64, device=self.dev)
        self.params[: self.M * self.D] = Zd.reshape(-1)
        o = self.M * self.D + self.M
        self.params[o:] = torch.eye(self.M, dtype=torch.float64, device=self.dev).reshape(-1)
        self.adam_m = torch.zeros_like(self.params)
        self.adam_v = torch.zeros_like(self.params)
        self.lr = float(lr)
        self.train_hyper = bool(train_hyper)
        self._tf = Softplus(0.0)
        self._u = self._tf.inverse(self.theta)
        self._hm = np.zeros(self.P)
        self._hv = np.zeros(self.P)
        self.t = 0
        self.cursor = 0
        import torch.distributed as dist
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1

    # views ------------------------------------------------------------------------------------
    @property
    def Z(self):
        return self.params[: self.M * self.D].view(self.M, self.D)

    @property
    def q_mu(self):
        return self.params[self.M * self.D: self.M * self.D + self.M]

    @property
    def q_sqrt(self):
        return self.params[self.M * self.D + self.M:].view(self.M, self.M)

    def _next_window(self):
        n = self.X.shape[0]
        if self.cursor + self.B > n:
            self.cursor = 0
        o = self.cursor
        self.cursor += self.B
        return o

    def step(self, update: bool = True) -> float:
        """One data-parallel ELBO + gradient evaluation (and optimiser update).  Returns the ELBO."""
        eng = self.engine
        ops.sync_stream(eng)
        eng.set_kernel(self.ck.spec, self.ck.token)
        o = self._next_window()
        Xb = self.X[o:o + self.B]
        yb = self.y[o:o + self.B]
        M, D, P = self.M, self.D, self.P
        eng.svgp_data_term(self.theta, self.noise, self.Z.data_ptr(), M, D, self.q_mu.data_ptr(), self.q_sqrt.data_ptr(), M,
                           Xb.data_ptr(), yb.data_ptr(), self.B, self.flat.data_ptr(), True)
        allreduce_sum_(self.flat, self.group)
        scale = float(self.num_data) / float(self.world * self.B)
        elbo, kl = eng.svgp_finish(self.flat.data_ptr(), scale, self.q_mu.data_ptr(), self.q_sqrt.data_ptr(), M, M, D, P, True)
        if update:
            self.t += 1
            g = self.flat[2 + P:]
            eng.adam_step(self.params.data_ptr(), g.data_ptr(), self.adam_m.data_ptr(), self.adam_v.data_ptr(), self.nvar,
                          self.lr, self.t, maximize=True)
            if self.train_hyper:
                gth = self.flat[2:2 + P].cpu().numpy() * self._tf.forward_grad(self._u)
                self._hm = 0.9 * self._hm + 0.1 * gth
                self._hv = 0.999 * self._hv + 0.001 * gth * gth
                mh = self._hm / (1 - 0.9 ** self.t)
                vh = self._hv / (1 - 0.999 ** self.t)
                self._u = self._u + self.lr * mh / (np.sqrt(vh) + 1e-8)
                self.theta = self._tf.forward(self._u)
        return elbo
