"""Prediction (ImportanceSample.extend + predictive_ll) without a GPU:
 (1) the oracle (oracle/predict_oracle.py) against the goldens of the UNMODIFIED reference run with explicit base noise
     (tests/golden/make_golden_predict.py);
 (2) the programs alan_b200.predict emits (VM draws, PasteOp, densities, reductions, LSE_eps), executed by the CPU
     emulator, against the same goldens."""
import os

import pytest
import torch as t

import models
from alan_b200 import model as M
from alan_b200.named import NT
from golden_io import GOLDEN_DIR, TAGS

CASES = ['cfg2_movielens', 'cfg3_radon']


def load(case, tag):
    return t.load(os.path.join(GOLDEN_DIR, f"predict_{case}_{tag}.pt"), weights_only=False)


def nts(d):
    return {k: NT(v[0], v[1]) for k, v in d.items()}


def prior_of(g, tag):
    P, _ = models.build(g['case'], M, TAGS[tag])
    return P


def close(a, b, tag, scale=1.0):
    tol = (3e-6 if tag == 'f32' else 1e-12) * scale
    return float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_predict_oracle_matches_reference(case, tag):
    from oracle import predict_oracle as PO
    g = load(case, tag)
    P = prior_of(g, tag)
    post, data = nts(g['post']), nts(g['data'])
    ext = PO.extend(P, post, data, nts(g['ext_inputs']), g['noise'], g['N'], TAGS[tag])
    assert set(ext) == set(g['extended'])
    for k, (ref, axes) in g['extended'].items():
        assert t.equal(ext[k].order(axes).t, ref), k
    latents = {k: v for k, v in ext.items() if k not in data}
    pll = PO.predictive_ll(P, latents, nts(g['ext_data']), g['platesizes'], nts(g['ext_inputs']), g['N'], TAGS[tag])
    for k, ref in g['pll'].items():
        assert close(pll[k], ref, tag), k


def emulate(prog, by_name, dtype):
    from plan_emulator import Emu
    ins = []
    for name in prog.plan.input_names:
        if name in prog.plan.const_inputs:
            ins.append(prog.plan.const_inputs[name])
        else:
            ins.append(by_name[name].to(dtype).contiguous())
    return ins


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_predict_programs_emulated_match_reference(case, tag):
    from alan_b200.predict import Extender, PredictiveLL
    from plan_emulator import Emu
    dt = TAGS[tag]
    g = load(case, tag)
    P = prior_of(g, tag)
    post, data, ext_in, ext_data = nts(g['post']), nts(g['data']), nts(g['ext_inputs']), nts(g['ext_data'])
    # ---- extend
    ex = Extender(P, post, data, g['ext_sizes'], ext_in, g['N'], dt)
    shapes = ex.noise_shapes()
    assert set(shapes) == set(g['noise'])
    for var, (kind, shape) in shapes.items():
        assert tuple(g['noise'][var].shape) == shape and g['noise_kinds'][var] == kind, var
    by_name = {name: g['noise'][var] for var, kind, axes, pos, name in ex.noise}
    src = {**data, **post}
    for name, k, axes in ex.orig_order:
        by_name[name] = src[k].order(axes).t
    for k, v in ext_in.items():
        by_name[k] = v.order(ex.in_axes[k]).t
    ins = emulate(ex, by_name, dt)
    numel = lambda axes, pos: max(1, int(t.tensor([ex.pl.sizes[a] for a in axes] + list(pos)).prod()))
    outs = {i: t.zeros(numel(axes, pos), dtype=dt) for i, (_, axes, pos) in enumerate(ex.outputs)}
    Emu(ex.plan, ins, outputs=outs).run(ex.plan.programs[0])
    ext = {}
    for i, (var, axes, pos) in enumerate(ex.outputs):
        ext[var] = NT(outs[i].reshape([ex.pl.sizes[a] for a in axes] + list(pos)), axes)
        ref, raxes = g['extended'][var]
        assert close(ext[var].order(raxes).t, ref, tag), var
    # ---- predictive log-likelihood on the reference's extended sample
    latents = {k: NT(*g['extended'][k]) for k in g['extended'] if k not in data}
    pl = PredictiveLL(P, latents, ext_data, g['platesizes'], ext_in, g['N'], dt)
    by_name = {k: v.order(pl.in_axes[k]).t for k, v in {**ext_in, **latents, **ext_data}.items()}
    ins = emulate(pl, by_name, dt)
    outs = {i: t.zeros(1, dtype=dt) for i in range(len(pl.vars))}
    Emu(pl.plan, ins, outputs=outs).run(pl.plan.programs[0])
    for i, var in enumerate(pl.vars):
        assert close(outs[i][0], g['pll'][var], tag, scale=3.0), var


def test_extend_errors_like_the_reference():
    from alan_b200.predict import Extender
    g = load('cfg2_movielens', 'f32')
    P = prior_of(g, 'f32')
    post, data, ext_in = nts(g['post']), nts(g['data']), nts(g['ext_inputs'])
    with pytest.raises(Exception, match="smaller than the original"):
        Extender(P, post, data, {**g['ext_sizes'], 'plate_1': 3}, {'x': NT(ext_in['x'].t[:3], ext_in['x'].axes)}, g['N'], t.float32)
