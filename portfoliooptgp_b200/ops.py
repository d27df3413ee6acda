"""Host<->device plumbing: torch owns device memory and streams, buffers reach the C-ABI as raw
pointers (zero-copy: ``data_ptr()`` of a contiguous fp64 CUDA tensor, or of a tensor imported
through DLPack).  Nothing here computes on the CPU."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _capi

_engines = {}


def cuda_device_index(device=None) -> int:
    if device is None:
        if not torch.cuda.is_available():
            raise _capi.EngineError("no CUDA device visible: portfoliooptgp_b200 has no CPU fallback")
        return torch.cuda.current_device()
    if isinstance(device, int):
        return device
    d = torch.device(device)
    return d.index if d.index is not None else torch.cuda.current_device()


def shared_engine(device=None) -> _capi.Engine:
    """A per-device engine for stateless helper calls (kernel.K(X), dense ops)."""
    idx = cuda_device_index(device)
    if idx not in _engines:
        _engines[idx] = _capi.Engine(idx)
    return _engines[idx]


def to_device(a, device=None, ndim: Optional[int] = None) -> torch.Tensor:
    """Anything array-like -> contiguous fp64 CUDA tensor.  CUDA tensors (torch, or any producer
    exposing ``__dlpack__``) are taken zero-copy; host arrays go through pinned memory."""
    idx = cuda_device_index(device)
    dev = torch.device("cuda", idx)
    if isinstance(a, torch.Tensor):
        t = a
    elif hasattr(a, "__dlpack__") and not isinstance(a, np.ndarray):
        t = torch.from_dlpack(a)
    else:
        arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
        t = torch.from_numpy(arr)
        if arr.nbytes >= (1 << 16):
            t = t.pin_memory()
    t = t.detach()
    if t.device != dev or t.dtype != torch.float64:
        t = t.to(device=dev, dtype=torch.float64, non_blocking=True)
    if ndim == 2 and t.ndim == 1:
        t = t[:, None]
    return t.contiguous()


def sync_stream(engine: _capi.Engine):
    """Point the engine at torch's current stream so engine work is ordered after torch's copies."""
    engine.set_stream(torch.cuda.current_stream(engine.device).cuda_stream)


def kernel_matrix(kernel, X, X2=None, *, mode: Optional[int] = None, diag_add: float = 0.0, device=None) -> torch.Tensor:
    """kernel(X, X2) on the GPU; returns a CUDA tensor [N, N2]."""
    from .kernels import compile_kernel
    Xd = to_device(X, device, ndim=2)
    X2d = None if X2 is None else to_device(X2, device, ndim=2)
    N, D = Xd.shape
    N2 = N if X2d is None else X2d.shape[0]
    ck = compile_kernel(kernel, D)
    eng = shared_engine(Xd.device.index)
    sync_stream(eng)
    eng.set_kernel(ck.spec)
    ld = (N2 + 1) // 2 * 2
    out = torch.empty((N, ld), dtype=torch.float64, device=Xd.device)
    if mode is None:
        mode = 2 if X2d is None else 0
    eng.assemble(ck.theta(), Xd.data_ptr(), N, None if X2d is None else X2d.data_ptr(), N2, D, out.data_ptr(), ld, mode,
                 diag_add)
    return out[:, :N2]


def kernel_diag(kernel, X, device=None) -> torch.Tensor:
    from .kernels import compile_kernel
    Xd = to_device(X, device, ndim=2)
    N, D = Xd.shape
    ck = compile_kernel(kernel, D)
    eng = shared_engine(Xd.device.index)
    sync_stream(eng)
    eng.set_kernel(ck.spec)
    out = torch.empty((N,), dtype=torch.float64, device=Xd.device)
    eng.kdiag(ck.theta(), Xd.data_ptr(), N, D, out.data_ptr())
    return out
