"""Data-parallel minibatch SVGP (north_star subsystem 5, BASELINE config C5).

One process per GPU.  Parameters are replicated; every rank owns a shard of the N rows
(pre-shuffled, so consecutive windows of the shard are uniform minibatches and are handed to the
engine zero-copy as pointer offsets), evaluates the UNSCALED data term and its gradient on its
minibatch, and the flat records are summed with ONE ``all_reduce`` (NCCL over NVLink; gloo in the
CPU tests) of the record with the q_sqrt gradient packed to its lower triangle (``PackedRecord``).  Then, identically on every rank: scale by num_data / (world * B), subtract the KL term
and its gradient once, and step the optimiser (Adam, as GPflow users do for
minibatch SVGP; the variational parameters and Z are updated on the device, the few constrained
hyper-parameters on the host).

Differences from a stock GPflow SVGP training loop, stated because a caller porting
test_scripts/SVGP.py:513-533 should know them: (1) the likelihood variance is FROZEN by default, as at
the reference call site (``set_trainable(model.likelihood.variance, False)``, SVGP.py:524); pass
``train_noise=True`` for GPflow's default (trainable, softplus + 1e-6 transform).  (2) minibatches are
consecutive windows of the local shard, which is re-permuted on the device at every epoch boundary, so
every row is drawn with the same probability even when the shard size is not a multiple of the minibatch
size (the n mod B rows left over at the end of one epoch's permutation are simply not used in that epoch).
(3) every rank must use the same minibatch size; this is checked at construction."""
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


class PackedRecord:
    """The flat SVGP record with the q_sqrt gradient reduced to its lower triangle for the all-reduce: the engine
    writes dS/dq_sqrt as a full [M, M] block of which only the lower triangle is meaningful (csrc/svgp.cu:23-24;
    ``svgp_finish`` zeroes the rest), so 2 + P + M D + M + M (M + 1) / 2 doubles cross the links instead of
    2 + P + M D + M + M^2 (M = 2048: 2.12 M instead of 4.21 M)."""

    def __init__(self, M: int, D: int, P: int, device):
        head = 2 + P + M * D + M
        tri = torch.tril_indices(M, M, device=device)
        self.index = torch.cat([torch.arange(head, device=device), head + tri[0] * M + tri[1]])
        self.buf = torch.empty(self.index.numel(), dtype=torch.float64, device=device)
        self.full_size = head + M * M

    def allreduce_(self, flat: torch.Tensor, group=None) -> torch.Tensor:
        if flat.numel() != self.full_size:
            raise ValueError("PackedRecord: record size does not match (M, D, P)")
        torch.index_select(flat, 0, self.index, out=self.buf)
        allreduce_sum_(self.buf, group)
        flat.index_copy_(0, self.index, self.buf)
        return flat


class ShardWindows:
    """Consecutive zero-copy minibatch windows over a shard that is re-permuted (in place, on its own device)
    at every epoch boundary.  ``next()`` returns the row offset of the next window of ``B`` rows."""

    def __init__(self, X: torch.Tensor, y: torch.Tensor, B: int, seed: int = 0, shuffle: bool = True):
        if B < 1 or B > X.shape[0]:
            raise ValueError("minibatch larger than the local shard")
        self.X, self.y, self.B, self.shuffle = X, y, int(B), bool(shuffle)
        self.gen = torch.Generator(device=X.device)
        self.gen.manual_seed(int(seed))
        self.cursor = 0
        self.epoch = 0

    def next(self) -> int:
        n = self.X.shape[0]
        if self.cursor + self.B > n:
            self.epoch += 1
            self.cursor = 0
            if self.shuffle:
                perm = torch.randperm(n, device=self.X.device, generator=self.gen)
                self.X.copy_(self.X[perm])
                self.y.copy_(self.y[perm])
        o = self.cursor
        self.cursor += self.B
        return o


def check_equal_minibatch(B: int, device, group=None) -> None:
    """Every rank scales its data term by num_data / (world * B): the minibatch size must be the same everywhere."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    t = torch.tensor([B, -B], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    if int(t[0]) != -int(t[1]):
        raise ValueError(f"SVGPDataParallel: minibatch sizes differ across ranks (min {-int(t[1])}, max {int(t[0])})")


def combine_records(S_sum: float, kl: float, num_data: float, world: int, B: int) -> float:
    """ELBO from the all-reduced data-term sum: scale once, subtract the KL once."""
    return float(num_data) / float(world * B) * S_sum - kl


class SVGPDataParallel:
    def __init__(self, kernel: Kernel, noise_variance: float, Z, num_data: int, X_shard, Y_shard, minibatch_size: int,
                 lr: float = 1e-2, train_hyper: bool = True, train_noise: bool = False, shuffle: bool = True, seed: int = 0,
                 device=None, group=None):
        self.device_index = ops.cuda_device_index(device)
        self.dev = torch.device("cuda", self.device_index)
        self.group = group
        self.X = ops.to_device(X_shard, self.device_index, ndim=2)
        Y = ops.to_device(Y_shard, self.device_index, ndim=2)
        self.y = Y[:, 0].contiguous()
        if shuffle:
            # the epoch permutation works in place: never on the caller's tensors (to_device is zero-copy for CUDA input)
            if isinstance(X_shard, torch.Tensor) and X_shard.is_cuda:
                self.X = self.X.clone()
            self.y = self.y.clone()
        self.D = int(self.X.shape[1])
        self.B = int(minibatch_size)
        if self.B > self.X.shape[0]:
            raise ValueError("minibatch larger than the local shard")
        check_equal_minibatch(self.B, self.dev, group)
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
        self.train_noise = bool(train_noise)
        self._ntf = Softplus(DEFAULT_VARIANCE_LOWER_BOUND)      # gpflow.likelihoods.Gaussian: softplus + 1e-6
        self._nu = float(self._ntf.inverse(np.asarray(self.noise)))
        self._nm = 0.0
        self._nv = 0.0
        self.t = 0
        import torch.distributed as dist
        rank = dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0
        self.windows = ShardWindows(self.X, self.y, self.B, seed=seed * 1000003 + rank, shuffle=shuffle)
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.packed = PackedRecord(self.M, self.D, self.P, self.dev) if self.world > 1 else None

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

    @property
    def cursor(self) -> int:
        return self.windows.cursor

    @cursor.setter
    def cursor(self, v: int):
        self.windows.cursor = int(v)

    def _next_window(self):
        return self.windows.next()

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
        if self.packed is not None:
            self.packed.allreduce_(self.flat, self.group)      # lower triangle of the q_sqrt gradient only
        scale = float(self.num_data) / float(self.world * self.B)
        elbo, kl = eng.svgp_finish(self.flat.data_ptr(), scale, self.q_mu.data_ptr(), self.q_sqrt.data_ptr(), M, M, D, P, True)
        if update:
            self.t += 1
            g = self.flat[2 + P:]
            eng.adam_step(self.params.data_ptr(), g.data_ptr(), self.adam_m.data_ptr(), self.adam_v.data_ptr(), self.nvar,
                          self.lr, self.t, maximize=True)
            host = self.flat[:2 + P].cpu().numpy() if (self.train_hyper or self.train_noise) else None
            if self.train_noise:
                # d ELBO / d noise_variance sits in flat[1] (scaled by svgp_finish); Adam on the unconstrained value
                gn = float(host[1]) * float(self._ntf.forward_grad(np.asarray(self._nu)))
                self._nm = 0.9 * self._nm + 0.1 * gn
                self._nv = 0.999 * self._nv + 0.001 * gn * gn
                self._nu += self.lr * (self._nm / (1 - 0.9 ** self.t)) / (np.sqrt(self._nv / (1 - 0.999 ** self.t)) + 1e-8)
                self.noise = float(self._ntf.forward(np.asarray(self._nu)))
            if self.train_hyper:
                gth = host[2:2 + P] * self._tf.forward_grad(self._u)
                self._hm = 0.9 * self._hm + 0.1 * gth
                self._hv = 0.999 * self._hv + 0.001 * gth * gth
                mh = self._hm / (1 - 0.9 ** self.t)
                vh = self._hv / (1 - 0.999 ** self.t)
                self._u = self._u + self.lr * mh / (np.sqrt(vh) + 1e-8)
                self.theta = self._tf.forward(self._u)
        return elbo
