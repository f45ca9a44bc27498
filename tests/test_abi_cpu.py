"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/alan_b200.h declares, parses a plan blob, and the product refuses to run without CUDA."""
import os
import re

import pytest
import torch as t

import models
from alan_b200 import model as M
from alan_b200 import runtime
from alan_b200.engine import Compiled
from golden_io import load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    runtime.build_library()
    L = runtime.lib()
    header = open(os.path.join(ROOT, "include", "alan_b200.h")).read()
    declared = set(re.findall(r"\b(alan_b200_[a-z0-9_]+)\s*\(", header))
    assert declared == set(runtime.EXPORTS)
    for sym in declared:
        assert hasattr(L, sym), sym
    assert L.alan_b200_abi_version() == 1


def test_plan_blob_round_trip_without_gpu(monkeypatch):
    import ctypes
    g = load("cfg2_movielens", "f32")
    P, Q = models.movielens_model(M)
    comp = Compiled(P, Q, g["sample_nt"], g["inputs_params_nt"], g["data_nt"], grad_names=list(g["params"]), N=4)
    blob = comp.plan.blob
    L = runtime.lib()
    h = ctypes.c_void_p()
    assert L.alan_b200_plan_create(ctypes.c_void_p(blob.data_ptr()), blob.numel(), ctypes.byref(h)) == 0
    assert L.alan_b200_num_inputs(h) == len(comp.plan.input_names)
    assert L.alan_b200_num_programs(h) == len(comp.plan.programs)
    assert L.alan_b200_workspace_bytes(h) == comp.plan.ws_bytes
    for i, prog in enumerate(comp.plan.programs):
        kernels = [op for op in prog if not type(op).__name__.startswith(('Fill', 'Deps'))]
        assert L.alan_b200_program_launches(h, i) == len(kernels)
    L.alan_b200_plan_destroy(h)
    # ALAN_B200_SEQ=1 (read when the plan is created): consecutive small ops run as one single-CTA launch
    monkeypatch.setenv("ALAN_B200_SEQ", "1")
    assert L.alan_b200_plan_create(ctypes.c_void_p(blob.data_ptr()), blob.numel(), ctypes.byref(h)) == 0
    fused = [L.alan_b200_program_launches(h, i) for i in range(len(comp.plan.programs))]
    for i, prog in enumerate(comp.plan.programs):
        assert 1 <= fused[i] <= len(prog)
    assert sum(fused[:2]) < sum(len(p_) for p_ in comp.plan.programs[:2]) // 2
    L.alan_b200_plan_destroy(h)
    monkeypatch.delenv("ALAN_B200_SEQ")
    bad = blob.clone()
    bad[0] = 0
    assert L.alan_b200_plan_create(ctypes.c_void_p(bad.data_ptr()), bad.numel(), ctypes.byref(h)) != 0
    assert b"magic" in L.alan_b200_last_error()


@pytest.mark.skipif(t.cuda.is_available(), reason="checks the no-GPU error path")
def test_no_cpu_fallback():
    from alan_b200.engine import Runner
    g = load("cfg1_lgl", "f32")
    P, Q = models.lgl_model(M)
    comp = Compiled(P, Q, g["sample_nt"], g["inputs_params_nt"], g["data_nt"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Runner(comp)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        runtime.lse_eps(t.zeros(3, 3))


def test_unsupported_lambda_raises():
    P = M.Plate(a=M.Normal(0., 1.), b=M.Normal(lambda a: a.cumsum(0), 1.))
    Q = M.Plate(a=M.Normal(0., 1.), b=M.Normal(0., 1.))
    from alan_b200.named import NT
    s = {'a': NT(t.randn(3), ('K_a',)), 'b': NT(t.randn(3), ('K_b',))}
    with pytest.raises(Exception, match="cannot trace"):
        Compiled(P, Q, s, {}, {})


def test_structure_errors_mirror_reference():
    with pytest.raises(Exception, match="duplicate names"):
        M.Plate(a=M.Normal(0., 1.), p=M.Plate(a=M.Normal(0., 1.)))
    with pytest.raises(Exception, match="Wrong number of arguments"):
        M.Normal(0.)
    P = M.Plate(a=M.Normal(0., 1.))
    Q = M.Plate(b=M.Normal(0., 1.))
    with pytest.raises(Exception, match="same variables"):
        M.check_PQ(P, Q, set())


def test_torch_custom_ops_are_registered_and_refuse_cpu_tensors():
    """The boundary north_star names: torch.ops.alan_b200.* over the C ABI (alan_b200/ops.py).  CUDA dispatch key
    only: handing them CPU tensors raises -- there is no CPU path behind the ops."""
    import torch
    from alan_b200 import ops
    for name in ops.OPS:
        assert hasattr(torch.ops.alan_b200, name), name
    schema = str(torch.ops.alan_b200.logpq_bwd.default._schema)
    assert "Tensor[] inputs" in schema and "grads" in schema and "ws" in schema
    with pytest.raises(NotImplementedError):
        torch.ops.alan_b200.gather(torch.zeros(2, 3, 4), torch.zeros(1, 2, dtype=torch.long), 2, 3, 4)
    with pytest.raises(NotImplementedError):
        torch.ops.alan_b200.logpq_fwd(0, 0, [torch.zeros(3)], torch.zeros(8, dtype=torch.uint8))
