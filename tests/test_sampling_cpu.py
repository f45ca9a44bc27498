"""Ancestral sampling of Q without a GPU:
 (1) the oracle (oracle/sample_oracle.py) against golden samples produced by the UNMODIFIED reference walk with
     explicit base noise (tests/golden/make_golden_sampling.py): bit-exact in both dtypes;
 (2) the sampling program alan_b200.sampling.QSampler emits (permutation / gather / factor-VM draw / Timeseries
     recursion ops), executed by the CPU emulator, against the same goldens."""
import os

import pytest
import torch as t

import models
from alan_b200 import model as M
from alan_b200.named import NT
from golden_io import GOLDEN_DIR, TAGS

CASES = ['cfg1_lgl', 'cfg1_lglp', 'cfg2_movielens', 'cfg3_radon', 'model1', 'ref_corr_q', 'cfg4_timeseries_P',
         'cfg1_lglp_indep', 'cfg3_radon_indep']          # *_indep: IndependentSampler (Problem.sample_nonmp)


def load(case, tag):
    return t.load(os.path.join(GOLDEN_DIR, f"qsample_{case}_{tag}.pt"), weights_only=False)


def plate_of(g):
    P, Q = models.build(g['case'], M, t.float64 if 'float64' in g['dtype'] else t.float32)
    return Q if g['side'] == 'Q' else P


def params_of(g):
    return {k: NT(v[0], v[1]) for d in (g['inputs'], g['params']) for k, v in d.items()}


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_sampling_oracle_matches_reference_walk(case, tag):
    from oracle.sample_oracle import sample_q
    g = load(case, tag)
    out = sample_q(plate_of(g), params_of(g), g['noise'], g['K'], g.get('sampler_mode', 0), TAGS[tag])
    assert set(out) == set(g['samples'])
    for k, (ref, axes) in g['samples'].items():
        assert t.equal(out[k].order(axes).t, ref), k


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_sampling_program_emulated_matches_reference_walk(case, tag):
    from alan_b200.sampling import QSampler, PermutationSampler, IndependentSampler
    from plan_emulator import Emu
    g = load(case, tag)
    ip = params_of(g)
    sampler = IndependentSampler if g.get('sampler_mode', 0) == 2 else PermutationSampler
    qs = QSampler(plate_of(g), ip, g['platesizes'], g['K'], sampler, TAGS[tag])
    shapes = qs.noise_shapes()
    assert set(shapes) <= set(g['noise'])
    for key, (kind, shape, dt) in shapes.items():
        assert tuple(g['noise'][key].shape) == shape and g['noise_kinds'][key] == kind, key
    ins = []
    by_input = {name: g['noise'][key] for key, kind, axes, pos, name in qs.noise}
    for name in qs.plan.input_names:
        if name in qs.plan.const_inputs:
            ins.append(qs.plan.const_inputs[name])
        elif name in by_input:
            ins.append(by_input[name].contiguous())
        else:
            axes = next(a for k, a in qs.param_order if k == name)
            ins.append(ip[name].order(axes).t.to(TAGS[tag]).contiguous())
    outs = {i: t.zeros(max(1, int(t.tensor([qs.pl.sizes[a] for a in axes] + list(pos)).prod())), dtype=TAGS[tag])
            for i, (_, axes, pos) in enumerate(qs.outputs)}
    emu = Emu(qs.plan, ins, outputs=outs)
    emu.run(qs.plan.programs[0])
    tol = 1e-6 if tag == 'f32' else 1e-13
    for i, (var, axes, pos) in enumerate(qs.outputs):
        ref, raxes = g['samples'][var]
        mine = NT(outs[i].reshape([qs.pl.sizes[a] for a in axes] + list(pos)), axes).order(raxes).t
        assert (mine - ref).abs().max() <= tol * max(1.0, float(ref.abs().max())), var


def test_unsupported_family_raises():
    from alan_b200.sampling import QSampler
    Q = M.Plate(p=M.Beta(1., 1.))
    with pytest.raises(Exception, match="not supported"):
        QSampler(Q, {}, {}, 4)


def mvn_sampling_case(case, tag, K=6, sampler=None):
    """QSampler of an MvN model + base noise + what the oracle draws from it (loc + scale_tril eps, torch's rsample)."""
    from alan_b200.sampling import QSampler, PermutationSampler
    from oracle.sample_oracle import sample_q
    dt = TAGS[tag]
    _, inputs_fn, kw = models.CASES[case][:3]
    inp = inputs_fn(**kw, dtype=dt)
    _, Q = models.build(case, M, dt)
    nt = lambda d: {k: NT(v.rename(None), tuple(n for n in v.names if n is not None)) for k, v in d.items()}
    ip = nt(inp['params'])
    qs = QSampler(Q, ip, inp['platesizes'], K, sampler or PermutationSampler, dt)
    g = t.Generator().manual_seed(21)
    noise = {}
    for key, (kind, shape, ndt) in qs.noise_shapes().items():
        noise[key] = (t.randn if kind == 'normal' else t.rand)(shape, dtype=t.float64, generator=g).to(ndt)
    want = sample_q(Q, ip, noise, K, 0, dt)
    return qs, ip, noise, want


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", ['mvn', 'mvn2'])
def test_mvn_sampling_program_emulated_vs_oracle(case, tag):
    from plan_emulator import Emu
    qs, ip, noise, want = mvn_sampling_case(case, tag)
    ins = []
    by_input = {name: noise[key] for key, kind, axes, pos, name in qs.noise}
    for name in qs.plan.input_names:
        if name in qs.plan.const_inputs:
            ins.append(qs.plan.const_inputs[name])
        elif name in by_input:
            ins.append(by_input[name].contiguous())
        else:
            axes = next(a for k, a in qs.param_order if k == name)
            ins.append(ip[name].order(axes).t.to(TAGS[tag]).contiguous())
    outs = {i: t.zeros(max(1, int(t.tensor([qs.pl.sizes[a] for a in axes] + list(pos)).prod())), dtype=TAGS[tag])
            for i, (_, axes, pos) in enumerate(qs.outputs)}
    Emu(qs.plan, ins, outputs=outs).run(qs.plan.programs[0])
    tol = 3e-6 if tag == 'f32' else 1e-12
    for i, (var, axes, pos) in enumerate(qs.outputs):
        mine = outs[i].reshape([qs.pl.sizes[a] for a in axes] + list(pos))
        ref = want[var].order(axes).t
        assert (mine - ref).abs().max() <= tol * max(1.0, float(ref.abs().max())), var
