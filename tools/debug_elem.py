#!/usr/bin/env python
"""Element-wise errors of marginals / moments of one golden case (engine vs f64 truth vs the reference's fp32)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch as t, models
from alan_b200 import model as M
from alan_b200.named import NT
from alan_b200.engine import Compiled, Runner
from golden_io import load, rel_err, elem_err, f64_truth, TAGS
from oracle import logpq_oracle as O
case, tag = (sys.argv[1] if len(sys.argv) > 1 else 'cfg2_movielens'), 'f32'
g = load(case, tag)
P, Q = models.build(case, M, TAGS[tag])
g2p, groups = Q.groupvarname2platenames(), Q.groupvarnames()
sizes = {**{a: s for v in g["sample_nt"].values() for a, s in v.named_sizes.items()}, **g["platesizes"]}
elf = {}
for key in g["marginals"]:
    gs = tuple(sorted(key, key=groups.index)); axes = tuple(M.Kname(x) for x in gs) + tuple(g2p[gs[0]])
    elf[key] = NT(t.zeros([sizes[a] for a in axes]), axes)
moms = [((v,), models.MOMENT_FUNCS[f]) for v, f in g["moment_specs"]]
tr = f64_truth(case, g, models, M, O, joints=[k for k in g["marginals"] if len(k) > 1])
for fast in (True, False):
    comp = Compiled(P, Q, g["sample_nt"], g["inputs_params_nt"], g["data_nt"], extra_log_factors=elf, moment_specs=moms,
                    grad_names=list(elf.keys()), fast_paths=fast)
    run = Runner(comp, "cuda:0")
    tens = run.device_inputs(g["sample_nt"], g["inputs_params_nt"], g["data_nt"], elf)
    lp = run.forward_raw(tens); grads = run.backward_raw(tens)
    print("fast", fast, "lp err", rel_err(lp.cpu(), g["elbo"]))
    for key, (ref, axes) in g["marginals"].items():
        name = comp.elf_keys[key]; pt = comp.plan.input_pts[name]
        mine = NT(grads[name].cpu(), pt.axes).order(axes).t
        tw = tr["marginals"][frozenset(key)]; truth = NT(tw.t, tw.axes).order(axes).t
        print("  ", key, "engine vs truth %.2e  ref vs truth %.2e  engine vs golden %.2e  sumK %s" % (
            elem_err(mine, truth), elem_err(ref, truth), elem_err(mine, ref), mine.sum().item() if mine.ndim == 1 else ''))
    for i, ((jname, plates, pos), (ref, axes)) in enumerate(zip(comp.moment_inputs, g["moments"])):
        mine = NT(grads[jname].cpu(), plates).order(axes).t if plates else grads[jname].cpu()
        truth = tr["moments"][i].order(axes).t
        print("  moment", jname, "engine vs truth %.2e  ref vs truth %.2e" % (elem_err(mine, truth), elem_err(ref, truth)))
