"""torch custom ops over the C ABI of libalan_b200.so  --  the boundary `north_star` names.

    torch.ops.alan_b200.logpq_fwd(plan, segment, inputs, ws)                  -> lp
    torch.ops.alan_b200.logpq_bwd(plan, segment, inputs, grad_lp, grads, ws)  -> ()      (writes grads)
    torch.ops.alan_b200.resample(plan, inputs, uniforms, idx_out, ws)         -> ()      (writes idx_out)
    torch.ops.alan_b200.run(plan, program, inputs, outputs, ws)               -> ()      (writes outputs)
    torch.ops.alan_b200.gather(x, idx, outer, K, inner)                       -> out

Each op is a `torch.library` registration (CUDA dispatch key only: a CPU tensor reaches no kernel and the
dispatcher raises -- there is no CPU path) whose body is one call of the matching `extern "C"` entry point of
include/alan_b200.h with raw device pointers, sizes and the caller's current CUDA stream.  `plan` is the address of
an `alan_b200_plan` (runtime.DevicePlan.handle).  The reference has no FFI layer: what these ops
replace is the Python call from `Sample._elbo` / `_importance_sample_idxs` / `index_into_sample` into
`logPQ_plate` / `logPQ_sample` (reference src/alan/Sample.py:92-106,159-177,359-381).  Autograd is attached one
level up (engine._LogPQFunction: forward = logpq_fwd segments [+ the tile all-reduce], backward = logpq_bwd).
"""
from __future__ import annotations

import ctypes
from typing import List

import torch

from . import runtime as _rt

_vp = ctypes.c_void_p


def _ptrs(tensors):
    arr = (_vp * max(len(tensors), 1))()
    for i, x in enumerate(tensors):
        arr[i] = x.data_ptr()
    return arr


def _stream(x: torch.Tensor):
    return _vp(torch.cuda.current_stream(x.device).cuda_stream)


@torch.library.custom_op("alan_b200::logpq_fwd", mutates_args=("ws",), device_types="cuda")
def logpq_fwd(plan: int, segment: int, inputs: List[torch.Tensor], ws: torch.Tensor) -> torch.Tensor:
    """alan_b200_logpq_fwd: one forward segment; returns the 0-d log-evidence tensor (written by the last segment)."""
    lp = torch.empty((), dtype=inputs[0].dtype if inputs else torch.float32, device=ws.device)
    with torch.cuda.device(ws.device):
        _rt.check(_rt.lib().alan_b200_logpq_fwd(_vp(plan), segment, _ptrs(inputs), _vp(lp.data_ptr()),
                                                _vp(ws.data_ptr()), _stream(ws)))
    return lp


@logpq_fwd.register_fake
def _(plan, segment, inputs, ws):
    return torch.empty((), dtype=inputs[0].dtype if inputs else torch.float32, device=ws.device)


@torch.library.custom_op("alan_b200::logpq_fwd_into", mutates_args=("lp", "ws"), device_types="cuda")
def logpq_fwd_into(plan: int, segment: int, inputs: List[torch.Tensor], lp: torch.Tensor, ws: torch.Tensor) -> None:
    """alan_b200_logpq_fwd writing into a caller-owned 0-d tensor (multi-segment forwards share one `lp`)."""
    with torch.cuda.device(ws.device):
        _rt.check(_rt.lib().alan_b200_logpq_fwd(_vp(plan), segment, _ptrs(inputs), _vp(lp.data_ptr()),
                                                _vp(ws.data_ptr()), _stream(ws)))


@torch.library.custom_op("alan_b200::logpq_bwd", mutates_args=("grads", "ws"), device_types="cuda")
def logpq_bwd(plan: int, segment: int, inputs: List[torch.Tensor], grad_lp: torch.Tensor, grads: List[torch.Tensor],
              ws: torch.Tensor) -> None:
    """alan_b200_logpq_bwd: one adjoint segment; `grads` (plan order) are written in place."""
    with torch.cuda.device(ws.device):
        _rt.check(_rt.lib().alan_b200_logpq_bwd(_vp(plan), segment, _ptrs(inputs), _vp(grad_lp.data_ptr()),
                                                _ptrs(grads), _vp(ws.data_ptr()), _stream(ws)))


@torch.library.custom_op("alan_b200::resample", mutates_args=("idx_out", "ws"), device_types="cuda")
def resample(plan: int, inputs: List[torch.Tensor], uniforms: List[torch.Tensor], idx_out: List[torch.Tensor],
             ws: torch.Tensor) -> None:
    """alan_b200_resample: posterior K indices from explicit float64 uniforms, written into `idx_out` (int64)."""
    with torch.cuda.device(ws.device):
        _rt.check(_rt.lib().alan_b200_resample(_vp(plan), _ptrs(inputs), _ptrs(uniforms), _ptrs(idx_out),
                                               _vp(ws.data_ptr()), _stream(ws)))


@torch.library.custom_op("alan_b200::run", mutates_args=("outputs", "ws"), device_types="cuda")
def run(plan: int, program: int, inputs: List[torch.Tensor], outputs: List[torch.Tensor], ws: torch.Tensor) -> None:
    """alan_b200_run: any program of a plan (stand-alone plans: Marginals.moments, Q sampling, prediction)."""
    with torch.cuda.device(ws.device):
        _rt.check(_rt.lib().alan_b200_run(_vp(plan), program, _ptrs(inputs), _ptrs(outputs), _vp(ws.data_ptr()),
                                          _stream(ws)))


@torch.library.custom_op("alan_b200::gather", mutates_args=(), device_types="cuda")
def gather(x: torch.Tensor, idx: torch.Tensor, outer: int, K: int, inner: int) -> torch.Tensor:
    """alan_b200_gather: x [outer, K, inner], idx [N, outer] int64 -> [N * outer * inner] (bit-exact copy)."""
    N = idx.shape[0]
    out = torch.empty(N * outer * inner, dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        _rt.check(_rt.lib().alan_b200_gather(_vp(x.data_ptr()), _vp(idx.data_ptr()), _vp(out.data_ptr()),
                                             x.element_size(), N, outer, K, inner, 1, _stream(x)))
    return out


@gather.register_fake
def _(x, idx, outer, K, inner):
    return torch.empty(idx.shape[0] * outer * inner, dtype=x.dtype, device=x.device)


OPS = ("logpq_fwd", "logpq_fwd_into", "logpq_bwd", "resample", "run", "gather")
