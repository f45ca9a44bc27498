"""Planner logic checked on CPU: plan ops are executed by tests/plan_emulator.py (torch
semantics of the CUDA kernels) and compared with the reference goldens.  The same plans run on
the GPU through the C ABI in tests/test_gpu_parity.py."""
import pytest
import torch as t

import models
from alan_b200 import model as M
from alan_b200.engine import Compiled
from alan_b200.named import NT
from golden_io import load, rel_err, tol, TAGS
from plan_emulator import Emu
from uniforms import UniformSource

CASES = list(models.CASES)


def run_fwd_bwd(comp, inputs, want_grads=True):
    plan = comp.plan
    lp = t.zeros(1, dtype=plan.dtype)
    emu = Emu(plan, inputs, outputs={0: lp})
    for seg in plan.programs[:plan.n_fwd]:
        emu.run(seg)
    grads = {}
    if want_grads and plan.n_bwd:
        emu.outputs = {i: t.zeros(plan.input_pts[n].numel, dtype=plan.dtype) for i, n in enumerate(plan.grad_inputs)}
        emu.aux = {0: t.ones(1, dtype=plan.dtype)}
        for seg in plan.programs[plan.n_fwd:plan.n_fwd + plan.n_bwd]:
            emu.run(seg)
        grads = {n: emu.outputs[i].reshape(plan.input_pts[n].shape) for i, n in enumerate(plan.grad_inputs)}
    return lp[0], grads, emu


def grad_as(comp, grads, name, axes):
    pt = comp.plan.input_pts[name]
    return NT(grads[name], pt.axes).order(axes).t if pt.axes else grads[name]


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_elbo_and_grads(case, tag):
    g = load(case, tag)
    P, Q = models.CASES[case][0](M)
    names = list(g["grad_sample"]) + list(g["grad_params"])
    comp = Compiled(P, Q, g["sample_nt"], g["inputs_params_nt"], g["data_nt"], grad_names=names)
    inputs = comp.canonical_inputs(g["sample_nt"], g["inputs_params_nt"], g["data_nt"])
    lp, grads, _ = run_fwd_bwd(comp, inputs)
    assert rel_err(lp, g["elbo"]) < tol(tag)
    for n in g["grad_sample"]:
        assert rel_err(grad_as(comp, grads, n, g["sample"][n][1]), g["grad_sample"][n]) < 30 * tol(tag), n
    for n in g["grad_params"]:
        assert rel_err(grad_as(comp, grads, n, g["params"][n][1]), g["grad_params"][n]) < 30 * tol(tag), n


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_marginals_and_moments(case, tag):
    g = load(case, tag)
    P, Q = models.CASES[case][0](M)
    dtype = TAGS[tag]
    g2p = Q.groupvarname2platenames()
    groups = Q.groupvarnames()
    elf = {}
    for key in g["marginals"]:
        gs = tuple(sorted(key, key=groups.index))
        axes = tuple(M.Kname(x) for x in gs) + tuple(g2p[gs[0]])
        sizes = {**{a: s for v in g["sample_nt"].values() for a, s in v.named_sizes.items()}, **g["platesizes"]}
        elf[key] = NT(t.zeros([sizes[a] for a in axes], dtype=dtype), axes)
    moms = [((v,), models.MOMENT_FUNCS[f]) for v, f in g["moment_specs"]]
    comp = Compiled(P, Q, g["sample_nt"], g["inputs_params_nt"], g["data_nt"], extra_log_factors=elf,
                    moment_specs=moms, grad_names=list(elf.keys()))
    inputs = comp.canonical_inputs(g["sample_nt"], g["inputs_params_nt"], g["data_nt"], elf)
    lp, grads, _ = run_fwd_bwd(comp, inputs)
    assert rel_err(lp, g["elbo"]) < tol(tag)
    for key, (ref, axes) in g["marginals"].items():
        name = comp.elf_keys[key]
        pt = comp.plan.input_pts[name]
        mine = NT(grads[name], pt.axes).order(axes).t
        assert rel_err(mine, ref) < 30 * tol(tag), key
    for (jname, plates, pos), (ref, axes) in zip(comp.moment_inputs, g["moments"]):
        mine = NT(grads[jname], plates).order(axes).t
        assert rel_err(mine, ref) < 30 * tol(tag), jname


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", [c for c in CASES if models.CASES[c][6] is not None])
def test_resampling(case, tag):
    g = load(case, tag)
    P, Q = models.CASES[case][0](M)
    N = g["N"]
    comp = Compiled(P, Q, g["sample_nt"], g["inputs_params_nt"], g["data_nt"], N=N)
    inputs = comp.canonical_inputs(g["sample_nt"], g["inputs_params_nt"], g["data_nt"])
    lp, _, emu = run_fwd_bwd(comp, inputs, want_grads=False)
    plan = comp.plan
    src = UniformSource(g["uniform_seed"], N, g["platesizes"], list(g["platesizes"]))
    aux = {}
    for i, (batch_axes, ks) in enumerate(plan.sample_steps):
        u, axes = src.draw(batch_axes)
        assert axes == tuple(batch_axes) + ('N',)
        aux[i] = u.reshape(-1)
    emu.aux = aux
    emu.outputs = {}
    for gi, (grp, plates) in enumerate(plan.sample_groups):
        n = N
        for a in plates:
            n *= g["platesizes"][a]
        emu.outputs[gi] = t.zeros(n, dtype=t.long)
    emu.run(plan.programs[plan.sample_prog])
    total = bad = 0
    for gi, (grp, plates) in enumerate(plan.sample_groups):
        ref, axes = g["indices"][grp]
        mine = NT(emu.outputs[gi].reshape([N] + [g["platesizes"][a] for a in plates]), ('N',) + tuple(plates))
        mine = mine.order(axes).t
        total += ref.numel()
        bad += (mine != ref).sum().item()
    assert bad <= 1e-3 * total, f"{bad}/{total}"
