"""Global importance sampling (`SampleNonMP`) on CPU: the oracle (oracle/nonmp_oracle.py)
against the goldens of the unmodified reference, and the non-MP plans (Planner.plan_nonmp) executed by the plan
emulator against the same goldens.  The GPU runs the same plans through the C ABI (tests/test_gpu_nonmp.py)."""
import os

import pytest
import torch as t

import models
from alan_b200 import model as M
from alan_b200.engine import Compiled
from alan_b200.named import NT
from alan_b200.nonmp import unify_K
from alan_b200.plan import NONMP_K
from golden_io import GOLDEN_DIR, TAGS, rel_err, tol
from oracle import nonmp_oracle as NO
from test_plan_emulated import run_fwd_bwd, grad_as

CASES = ['cfg1_lgl', 'cfg1_lglp', 'cfg2_movielens', 'cfg3_radon', 'model1', 'ref_bernoulli']


def load(case, tag):
    g = t.load(os.path.join(GOLDEN_DIR, f"nonmp_{case}_{tag}.pt"), weights_only=False)
    nt = lambda d: {k: NT(v[0], v[1]) for k, v in d.items()}
    g["sample_nt"] = nt(g["sample"])
    g["inputs_params_nt"] = {**nt(g["inputs"]), **nt(g["params"])}
    g["data_nt"] = nt(g["data"])
    return g


def unified_axes(axes):
    return tuple(NONMP_K if a.startswith('K_') else a for a in axes)


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_oracle_vs_reference_golden(case, tag):
    g = load(case, tag)
    P, Q = models.build(case, M, TAGS[tag])
    args = (P, Q, g["sample_nt"], g["inputs_params_nt"], g["data_nt"])
    lpq = NO.logpq(*args).order(('K',)).t
    assert rel_err(lpq, g["lpq"]) < tol(tag)
    assert rel_err(NO.elbo(*args), g["elbo"]) < tol(tag)
    moms = [((v,), models.MOMENT_FUNCS[f]) for v, f in g["moment_specs"]]
    for mine, (ref, axes) in zip(NO.moments(*args, moms), g["moments"]):
        assert rel_err(mine.order(axes).t, ref) < 30 * tol(tag)
    # gradients of the oracle's elbo by autograd
    leaves = {k: NT(v.t.clone().requires_grad_(), v.axes) for k, v in g["sample_nt"].items()}
    ip = {k: NT(v.t.clone().requires_grad_(k in g["grad_params"]), v.axes) for k, v in g["inputs_params_nt"].items()}
    L = NO.elbo(P, Q, leaves, ip, g["data_nt"])
    names = list(g["grad_sample"]) + list(g["grad_params"])
    grads = t.autograd.grad(L, [leaves[n].t for n in g["grad_sample"]] + [ip[n].t for n in g["grad_params"]])
    for n, gr in zip(names, grads):
        ref = g["grad_sample"][n] if n in g["grad_sample"] else g["grad_params"][n]
        assert rel_err(gr, ref) < 30 * tol(tag), n


def test_oracle_inverse_cdf_draw_follows_the_weights():
    g = load('cfg1_lgl', 'f64')
    P, Q = models.build('cfg1_lgl', M, t.float64)
    args = (P, Q, g["sample_nt"], g["inputs_params_nt"], g["data_nt"])
    w = NO.weights(*args)
    u = (t.arange(4000, dtype=t.float64) + 0.5) / 4000
    idx = NO.importance_sample_idxs(*args, u)
    freq = t.bincount(idx, minlength=w.numel()).double() / 4000
    assert float((freq - w).abs().max()) < 1e-3


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_emulated_plan_vs_reference_golden(case, tag):
    g = load(case, tag)
    P, Q = models.build(case, M, TAGS[tag])
    sample = unify_K(g["sample_nt"])
    names = list(g["grad_sample"]) + list(g["grad_params"])
    comp = Compiled(P, Q, sample, g["inputs_params_nt"], g["data_nt"], grad_names=names, nonmp=True)
    inputs = comp.canonical_inputs(sample, g["inputs_params_nt"], g["data_nt"])
    lp, grads, _ = run_fwd_bwd(comp, inputs)
    assert rel_err(lp, g["elbo"]) < tol(tag)
    for n in g["grad_sample"]:
        assert rel_err(grad_as(comp, grads, n, unified_axes(g["sample"][n][1])), g["grad_sample"][n]) < 30 * tol(tag), n
    for n in g["grad_params"]:
        assert rel_err(grad_as(comp, grads, n, g["params"][n][1]), g["grad_params"][n]) < 30 * tol(tag), n
    # moments: gradient with respect to the zero source terms
    moms = [((v,), models.MOMENT_FUNCS[f]) for v, f in g["moment_specs"]]
    comp = Compiled(P, Q, sample, g["inputs_params_nt"], g["data_nt"], moment_specs=moms, nonmp=True)
    inputs = comp.canonical_inputs(sample, g["inputs_params_nt"], g["data_nt"])
    lp, grads, _ = run_fwd_bwd(comp, inputs)
    for (jname, plates, pos), (ref, axes) in zip(comp.moment_inputs, g["moments"]):
        assert rel_err(NT(grads[jname], plates).order(axes).t, ref) < 30 * tol(tag), jname


@pytest.mark.parametrize("case", ['cfg1_lglp', 'cfg3_radon'])
def test_emulated_resampling_vs_oracle(case):
    tag = 'f64'
    g = load(case, tag)
    P, Q = models.build(case, M, TAGS[tag])
    sample = unify_K(g["sample_nt"])
    N = 64
    comp = Compiled(P, Q, sample, g["inputs_params_nt"], g["data_nt"], N=N, nonmp=True)
    inputs = comp.canonical_inputs(sample, g["inputs_params_nt"], g["data_nt"])
    lp, _, emu = run_fwd_bwd(comp, inputs, want_grads=False)
    u = t.rand(N, dtype=t.float64, generator=t.Generator().manual_seed(5))
    emu.aux = {0: u}
    emu.outputs = {0: t.zeros(N, dtype=t.long)}
    emu.run(comp.plan.programs[comp.plan.sample_prog])
    ref = NO.importance_sample_idxs(P, Q, g["sample_nt"], g["inputs_params_nt"], g["data_nt"], u)
    assert t.equal(emu.outputs[0], ref)


def test_timeseries_is_refused_like_the_reference():
    g = t.load(os.path.join(GOLDEN_DIR, "cfg4_timeseries_f32.pt"), weights_only=False)
    nt = lambda d: {k: NT(v[0], v[1]) for k, v in d.items()}
    P, Q = models.build('cfg4_timeseries', M, t.float32)
    K = g['K']
    sample = {'init': NT(t.randn(K), (NONMP_K,)), 'ts': NT(t.randn(K, 37), (NONMP_K, 'T'))}
    with pytest.raises(Exception, match="Timeseries"):
        Compiled(P, Q, sample, {**nt(g['inputs']), **nt(g['params'])}, nt(g['data']), nonmp=True)
