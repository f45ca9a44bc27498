"""alan_b200.alan_adapter against the LIVE reference (build container only; skipped where
/root/reference is absent).  The reference builds the model, samples Q and evaluates its own
`elbo_vi`; the adapter converts the very same objects (Plate trees, torchdim tensors) and the plan is
executed by tests/plan_emulator.py -- the value must agree with the reference's within 1e-5 (fp32)."""
import pytest
import torch as t

import models
from golden_io import rel_err
from oracle.refcompat import reference_available

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container only)")]


def _problem(alan, case, seed=0, dtype=t.float32):
    model, inputs_fn, kw, K, moms, joints, N = models.CASES[case]
    t.manual_seed(seed)
    inp = inputs_fn(**kw, seed=seed, dtype=dtype)
    P, Q = model(alan)
    bp = alan.BoundPlate(P, inp['platesizes'], inputs=inp['inputs'])
    bq = alan.BoundPlate(Q, inp['platesizes'], inputs=inp['inputs'],
                         extra_opt_params={k: v.clone() for k, v in inp['params'].items()})
    return alan.Problem(bp, bq, inp['data']), K


@pytest.mark.parametrize("case", list(models.CASES))
def test_adapter_matches_reference_elbo(case):
    from oracle.refcompat import import_reference
    from alan_b200 import alan_adapter as A
    from alan_b200.engine import Compiled
    from test_plan_emulated import run_fwd_bwd
    alan = import_reference()
    prob, K = _problem(alan, case)
    s = prob.sample(K, reparam=False)
    ref = s.elbo_rws(computation_strategy=alan.no_checkpoint).detach()
    Pm, Qm = A.plate_from_reference(prob.P.plate), A.plate_from_reference(prob.Q.plate)
    sample = A.nts_from_tree(s.detached_sample)
    ip = A.nts_from_tree(prob.inputs_params())
    data = A.nts_from_tree(prob.data)
    comp = Compiled(Pm, Qm, sample, {k: v.detach() for k, v in ip.items()}, data)
    inputs = comp.canonical_inputs(sample, ip, data)
    lp, _, _ = run_fwd_bwd(comp, inputs, want_grads=False)
    assert rel_err(lp, ref) < 1e-5


def test_adapter_model_tree_roundtrip():
    """The converted tree has the same groups, plates and K axes as the reference's."""
    from oracle.refcompat import import_reference
    from alan_b200 import alan_adapter as A
    alan = import_reference()
    for case in models.CASES:
        prob, K = _problem(alan, case)
        Qm = A.plate_from_reference(prob.Q.plate)
        mirror = models.CASES[case][0](__import__("alan_b200.model", fromlist=["x"]))[1]
        assert Qm.groupvarnames() == mirror.groupvarnames()
        assert Qm.all_platenames() == mirror.all_platenames()
        assert Qm.varname2groupvarname() == mirror.varname2groupvarname()


@pytest.mark.parametrize("case", ["cfg1_lglp", "cfg2_movielens", "cfg3_radon", "model1"])
def test_b200_strategy_hook_on_live_reference_objects(case, monkeypatch):
    """`alan_adapter.B200` + `install_hook` from a live reference `Problem.sample(K)`: log-evidence, parameter
    gradients (through autograd into the reference's own named parameters), marginals and moments, with the plan
    executed by the CPU emulator (tests/plan_emulator.EmuRunner stands in for engine.Runner here; the same test runs
    the CUDA engine on the GPU box: tests/test_reference_gpu.py)."""
    from oracle.refcompat import import_reference
    from alan_b200 import alan_adapter as A, engine
    from plan_emulator import EmuRunner
    alan = import_reference()
    from alan.utils import generic_dims, generic_order
    monkeypatch.setattr(engine, "Runner", EmuRunner)
    prob, K = _problem(alan, case)
    moms, joints = models.CASES[case][4], models.CASES[case][5]
    s = prob.sample(K, reparam=False)
    params = dict(prob.Q._opt_params.to_dict())
    ref = s.elbo_rws(computation_strategy=alan.no_checkpoint)
    ref_g = t.autograd.grad(ref, list(params.values()), allow_unused=True)
    ref_marg = s.marginals(joints=joints, computation_strategy=alan.no_checkpoint)
    mlist = [((v,), alan.moments.RawMoment(models.MOMENT_FUNCS[f])) for v, f in moms]
    ref_mom = s._moments_uniform_input(mlist)
    remove = A.install_hook(alan)
    try:
        strat = A.B200()
        L = s.elbo_rws(computation_strategy=strat)
        assert rel_err(L.detach(), ref.detach()) < 1e-5
        g = t.autograd.grad(L, list(params.values()), allow_unused=True)
        for (n, _), a, b in zip(params.items(), g, ref_g):
            if b is not None:
                assert rel_err(a.rename(None), b.rename(None)) < 3e-4, n
        marg = s.marginals(joints=joints, computation_strategy=strat)
        for key, w in ref_marg.weights.items():
            dims = generic_dims(w)
            assert rel_err(generic_order(marg.weights[key], dims), generic_order(w, dims)) < 3e-4, key
        mom = s._moments_uniform_input(mlist, computation_strategy=strat)
        for a, b in zip(mom, ref_mom):
            dims = generic_dims(b)
            assert rel_err(generic_order(a, dims), generic_order(b, dims)) < 3e-4
    finally:
        remove()
