"""ctypes binding of libalan_b200.so and the torch-side plumbing around it.

PyTorch is used here for device memory, streams and autograd bookkeeping only; every
arithmetic step of the logPQ path runs in the hand-written kernels behind the C ABI
(include/alan_b200.h).  The library is built in-tree by `__graft_entry__.build()`; if it is
missing, or no CUDA device is present, every entry point raises -- there is no CPU path.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ALAN_B200_LIB") or os.path.join(_HERE, "libalan_b200.so")     # override: instrumented builds
SRC = os.path.join(_HERE, "csrc", "alan_b200.cu")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-split-compile", "0",          # one translation unit, optimised in parallel: 4 min -> 1.7 min on 8 cores
              "-shared", "-Xcompiler", "-fPIC"]

EXPORTS = [
    "alan_b200_abi_version", "alan_b200_last_error", "alan_b200_plan_create", "alan_b200_plan_destroy",
    "alan_b200_workspace_bytes", "alan_b200_num_inputs", "alan_b200_num_programs",
    "alan_b200_program_launches", "alan_b200_run", "alan_b200_profile", "alan_b200_logpq_fwd", "alan_b200_logpq_bwd",
    "alan_b200_resample", "alan_b200_gather", "alan_b200_lse_eps", "alan_b200_chain_scratch_elems",
    "alan_b200_logmmexp_chain", "alan_b200_normal_logpdf_bcast", "alan_b200_pipe_peak",
    "alan_b200_comm_bytes", "alan_b200_plan_set_comm", "alan_b200_qem_update", "alan_b200_widen_u8",
]


def _source_digest():
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc")))
    files.append(os.path.join(os.path.dirname(_HERE), "include", "alan_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> alan_b200/libalan_b200.so.  Skipped only when the
    library exists AND was built from exactly these sources and flags (content hash in libalan_b200.so.src)."""
    digest, stamp = _source_digest(), LIB_PATH + ".src"
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        if verbose:
            print(f"{LIB_PATH} is up to date with its sources ({digest[:12]})")
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB_PATH, SRC]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    open(stamp, "w").write(digest)
    return LIB_PATH


_lib = None


def _register_ops():
    from . import ops  # noqa: F401  (registers torch.ops.alan_b200.*)


def lib():
    """The loaded C-ABI library.  Fails loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the B200 engine has no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    L.alan_b200_abi_version.restype = i32
    L.alan_b200_last_error.restype = ctypes.c_char_p
    L.alan_b200_plan_create.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(vp)]
    L.alan_b200_plan_destroy.argtypes = [vp]
    L.alan_b200_workspace_bytes.argtypes = [vp]
    L.alan_b200_workspace_bytes.restype = ctypes.c_size_t
    L.alan_b200_comm_bytes.argtypes = [vp]
    L.alan_b200_comm_bytes.restype = ctypes.c_size_t
    L.alan_b200_plan_set_comm.argtypes = [vp, i32, i32, vp, ctypes.c_size_t]
    L.alan_b200_num_inputs.argtypes = [vp]
    L.alan_b200_num_programs.argtypes = [vp]
    L.alan_b200_program_launches.argtypes = [vp, i32]
    L.alan_b200_run.argtypes = [vp, i32, vp, vp, vp, vp]
    L.alan_b200_profile.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, i32]
    L.alan_b200_logpq_fwd.argtypes = [vp, i32, vp, vp, vp, vp]
    L.alan_b200_logpq_bwd.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.alan_b200_resample.argtypes = [vp, vp, vp, vp, vp, vp]
    L.alan_b200_gather.argtypes = [vp, vp, vp, i32, i64, i64, i64, i64, i64, vp]
    L.alan_b200_lse_eps.argtypes = [vp, vp, i64, i64, i32, vp]
    L.alan_b200_chain_scratch_elems.argtypes = [i64, i64, i64]
    L.alan_b200_chain_scratch_elems.restype = i64
    L.alan_b200_logmmexp_chain.argtypes = [vp, vp, vp, i64, i64, i64, i32, vp]
    L.alan_b200_normal_logpdf_bcast.argtypes = [vp, vp, vp, vp, i64, i64, vp, vp, vp, i32, vp]
    L.alan_b200_qem_update.argtypes = [i32, i64, ctypes.c_double, vp, vp, vp, vp, vp, vp, i32, vp]
    L.alan_b200_widen_u8.argtypes = [vp, vp, i64, i32, vp]
    L.alan_b200_pipe_peak.argtypes = [i32, vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_double), vp]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise Exception("alan_b200: " + lib().alan_b200_last_error().decode())


def require_cuda(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("alan_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device(device if device is not None else "cuda")


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * max(len(tensors), 1))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


def _stream(device=None):
    """The caller's current stream ON `device` (not on whatever device happens to be current)."""
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class DevicePlan:
    """A plan handle plus its device workspace."""
    def __init__(self, plan, device):
        self.plan = plan
        self.device = require_cuda(device)
        L = lib()
        blob = plan.blob.contiguous()
        self._blob = blob
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(L.alan_b200_plan_create(ctypes.c_void_p(blob.data_ptr()), blob.numel(), ctypes.byref(h)))
        self.handle = h
        nbytes = L.alan_b200_workspace_bytes(h)
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.consts = {k: v.to(self.device) for k, v in plan.const_inputs.items()}
        self.launches = [L.alan_b200_program_launches(h, i) for i in range(L.alan_b200_num_programs(h))]

    def attach_symmetric(self, process_group):
        """Allocate this plan's symmetric buffer (torch.distributed._symmetric_memory: cuMem allocations exchanged
        between the ranks of one node and mapped into every rank's address space over NVLink), rendezvous, and hand
        the peer mappings to the plan (alan_b200_plan_set_comm).  Collective: every rank of the group calls it."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        L = lib()
        nbytes = int(L.alan_b200_comm_bytes(self.handle))
        if nbytes == 0:
            return
        pg = process_group if process_group is not None else dist.group.WORLD
        with torch.cuda.device(self.device):
            buf = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
            buf.zero_()
            hdl = symm.rendezvous(buf, pg.group_name if hasattr(pg, "group_name") else pg)
            torch.cuda.synchronize(self.device)
            dist.barrier(group=process_group)                      # every rank's buffer is zeroed before anyone signals
            ptrs = (ctypes.c_void_p * hdl.world_size)(*[int(p) for p in hdl.buffer_ptrs])
            check(L.alan_b200_plan_set_comm(self.handle, hdl.rank, hdl.world_size, ptrs, nbytes))
        self._symm = (buf, hdl)

    def __del__(self):
        try:
            if self.handle:
                lib().alan_b200_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # The typed entry points go through the torch custom-op layer (alan_b200/ops.py: torch.ops.alan_b200.*), which
    # enters the plan's device and passes raw pointers + the caller's current stream on that device to the C ABI.
    def fwd(self, segment, inputs, lp_out):
        torch.ops.alan_b200.logpq_fwd_into(self.handle.value, segment, list(inputs), lp_out, self.ws)

    def bwd(self, segment, inputs, grad_lp, grads_out):
        torch.ops.alan_b200.logpq_bwd(self.handle.value, segment, list(inputs), grad_lp, list(grads_out), self.ws)

    def resample(self, inputs, uniforms, idx_out):
        torch.ops.alan_b200.resample(self.handle.value, list(inputs), list(uniforms), list(idx_out), self.ws)

    def run(self, program, inputs, outputs):
        """Generic program run (alan_b200_run): used by the stand-alone plans (Marginals.moments, Q sampling)."""
        torch.ops.alan_b200.run(self.handle.value, program, list(inputs), list(outputs), self.ws)

    def profile(self, program, inputs, outputs, aux):
        """per-op device milliseconds of one program run (CUDA events around every op)."""
        n = len(self.plan.programs[program])
        ms = (ctypes.c_float * max(n, 1))()
        with torch.cuda.device(self.device):
            got = lib().alan_b200_profile(self.handle, program, _ptr_array(inputs), _ptr_array(outputs),
                                          _ptr_array(aux), ctypes.c_void_p(self.ws.data_ptr()), _stream(self.device), ms, n)
        if got < 0:
            raise Exception("alan_b200: " + lib().alan_b200_last_error().decode())
        return [ms[i] for i in range(got)]

    def ws_view(self, pt, dtype):
        """torch view of a workspace tensor (used for the cross-GPU all-reduce of the plate tile)."""
        item = torch.empty((), dtype=dtype).element_size()
        return self.ws[pt.offset: pt.offset + pt.numel * item].view(dtype).view(pt.shape if pt.shape else ())


# ---- unit-level ops (parity tests call these through the C ABI) --------------------------------

def _dt(t):
    if t.dtype == torch.float32:
        return 0
    if t.dtype == torch.float64:
        return 1
    raise Exception("alan_b200 kernels compute in float32 or float64")


def lse_eps(x: torch.Tensor) -> torch.Tensor:
    """log(sum_r exp(x - max) + eps) + max over the last dim (reference utils.py:207-222)."""
    require_cuda()
    x = x.contiguous()
    if x.numel() == 0:
        raise Exception("lse_eps: empty input")
    n_red = x.shape[-1]
    out = torch.empty(x.shape[:-1], dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().alan_b200_lse_eps(x.data_ptr(), out.data_ptr(), out.numel(), n_red, _dt(x), _stream(x.device)))
    return out


def logmmexp_chain(ms: torch.Tensor) -> torch.Tensor:
    """ms [outer, T, K, K] -> logsumexp_Kcurr(chain_logmmexp(ms)) [outer, K] (utils.py:478-510, logpq.py:134-143)."""
    require_cuda()
    ms = ms.contiguous()
    outer, T, K, K2 = ms.shape
    assert K == K2
    L = lib()
    levels = torch.empty(L.alan_b200_chain_scratch_elems(outer, T, K), dtype=ms.dtype, device=ms.device)
    out = torch.empty(outer, K, dtype=ms.dtype, device=ms.device)
    with torch.cuda.device(ms.device):
        check(L.alan_b200_logmmexp_chain(ms.data_ptr(), levels.data_ptr(), out.data_ptr(), outer, T, K, _dt(ms),
                                         _stream(ms.device)))
    return out


def normal_logpdf_bcast(value, loc, scale, n_cells, n_event, vs, ls, ss) -> torch.Tensor:
    """sum_e Normal(loc, scale).log_prob(value) over a broadcast [n_cells, n_event] space."""
    require_cuda()
    out = torch.empty(n_cells, dtype=value.dtype, device=value.device)
    mk = lambda s: (ctypes.c_int64 * 2)(*s)
    with torch.cuda.device(value.device):
        check(lib().alan_b200_normal_logpdf_bcast(value.data_ptr(), loc.data_ptr(), scale.data_ptr(), out.data_ptr(),
                                                  n_cells, n_event, mk(vs), mk(ls), mk(ss), _dt(value),
                                                  _stream(value.device)))
    return out


def pipe_peak(which: int, device=None) -> float:
    """Measured lane-operations per second of one SM pipe (0: MUFU.EX2, 1: FFMA) on `device` -- bench.py's
    roofline denominators for the MUFU- and FP32-bound kernels, taken in the same process as the bench."""
    dev = require_cuda(device)
    scratch = torch.empty(4 * 1024 * 1024, dtype=torch.float32, device=dev)
    r = ctypes.c_double(0.0)
    with torch.cuda.device(dev):
        check(lib().alan_b200_pipe_peak(which, scratch.data_ptr(), scratch.numel() * 4, ctypes.byref(r), _stream(dev)))
    return r.value


QEM_FAMILY = {"Normal": 0, "Bernoulli": 1, "Poisson": 2, "Exponential": 3, "HalfNormal": 4, "Gamma": 5, "Beta": 6}


def qem_update(family: str, lr: float, new, means, params):
    """In place: means <- means * (1 - lr) + lr * new, params <- mean2conv(means) for one latent variable
    (reference BoundPlate.py:256-296, conversions.py:46-296).  `new`, `means`, `params`: lists of same-shape
    contiguous device tensors (1 or 2 each, per family)."""
    require_cuda()
    if family not in QEM_FAMILY:
        raise Exception(f"QEM: no mean <-> conventional parameter conversion for {family} on the device")
    ts = list(new) + list(means) + list(params)
    for x in ts:
        if not x.is_cuda or not x.is_contiguous() or x.dtype != ts[0].dtype or x.numel() != ts[0].numel():
            raise Exception("qem_update: moments, means and parameters must be contiguous device tensors of one shape and dtype")
    ptr = lambda xs, i: xs[i].data_ptr() if i < len(xs) else None
    x0 = ts[0]
    with torch.cuda.device(x0.device):
        check(lib().alan_b200_qem_update(QEM_FAMILY[family], x0.numel(), float(lr), ptr(new, 0), ptr(new, 1), ptr(means, 0),
                                         ptr(means, 1), ptr(params, 0), ptr(params, 1), _dt(x0), _stream(x0.device)))


NARROW_DTYPES = (torch.uint8, torch.bool)


def widen(src: torch.Tensor, dst: torch.Tensor = None, dtype=torch.float32) -> torch.Tensor:
    """uint8 / bool device tensor -> working dtype (alan_b200_widen_u8); `dst` may be a preallocated buffer."""
    require_cuda()
    if src.dtype not in NARROW_DTYPES or not src.is_cuda or not src.is_contiguous():
        raise Exception("widen: source must be a contiguous uint8 / bool device tensor")
    if dst is None:
        dst = torch.empty(src.shape, dtype=dtype, device=src.device)
    if dst.numel() != src.numel() or not dst.is_contiguous() or dst.device != src.device:
        raise Exception("widen: destination must be a contiguous tensor of the same size on the same device")
    with torch.cuda.device(src.device):
        check(lib().alan_b200_widen_u8(src.data_ptr(), dst.data_ptr(), src.numel(), _dt(dst), _stream(src.device)))
    return dst


def gather(x: torch.Tensor, idx: torch.Tensor, outer: int, K: int, inner: int) -> torch.Tensor:
    """x [outer, K, inner], idx [N, outer] int64 -> out [N, outer, inner] (Sample.py:359-381)."""
    require_cuda()
    return torch.ops.alan_b200.gather(x.contiguous(), idx.contiguous(), outer, K, inner)


_register_ops()
