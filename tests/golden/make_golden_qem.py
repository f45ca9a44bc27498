"""Golden vectors of the reference's QEM update, from the UNMODIFIED reference.

    python tests/golden/make_golden_qem.py              # writes tests/golden/qem_<case>_<dtype>.pt

Per case of tests/models.py::QEM_CASES: bind the model (`BoundPlate`: OptParam / QEMParam -> named parameters,
initial mean parameters), draw ONE sample with `problem.sample(K)`, then call `sample.update_qem_params(lr)` once per
listed learning rate on that same sample (Sample.py:351-355 -> BoundPlate.py:256-296 -> conversions.py) and record
the QEM parameters and moving-average means of P and Q after every call.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch as t
from oracle.refcompat import import_reference

alan = import_reference()
from alan.utils import generic_dims, generic_order
from alan.Plate import flatten_tree

import models
from make_golden import named_plain


def snap(prob):
    out = {}
    for side, bp in (('P', prob.P), ('Q', prob.Q)):
        out[side] = {'params': {k: named_plain(v) for k, v in bp.qem_params().items()},
                     'means': {k: named_plain(v) for k, v in bp.qem_means().items()}}
    return out


def run_case(name, dtype, seed=0):
    model, inputs_fn, kw, K, lrs = models.QEM_CASES[name]
    t.set_default_dtype(dtype)
    t.manual_seed(seed + 23)
    inp = inputs_fn(**kw, seed=seed, dtype=dtype)
    P, Q = model(alan)
    bp = alan.BoundPlate(P, inp['platesizes'], inputs=inp['inputs'])
    bq = alan.BoundPlate(Q, inp['platesizes'], inputs=inp['inputs'],
                         extra_opt_params={k: v.clone() for k, v in inp['params'].items()})
    prob = alan.Problem(bp, bq, inp['data'])
    s = prob.sample(K, reparam=False)
    flat = flatten_tree(s.detached_sample)
    plate_order = list(inp['platesizes'])
    sample = {}
    for k, v in flat.items():
        dims = generic_dims(v)
        kd = [d for d in dims if str(d).startswith('K_')]
        pl = [d for n in plate_order for d in dims if str(d) == n]
        sample[k] = (generic_order(v, [*kd, *pl]).detach().clone(), tuple(str(d) for d in [*kd, *pl]))
    out = {
        'case': name, 'K': K, 'dtype': str(dtype), 'platesizes': inp['platesizes'], 'lrs': lrs,
        'sample': sample,
        'params': {k: named_plain(v) for k, v in inp['params'].items()},
        'opt_params': {k: named_plain(v) for k, v in prob.Q.opt_params().items()},
        'data': {k: named_plain(v) for k, v in inp['data'].items()},
        'states': [snap(prob)],
    }
    for lr in lrs:
        s.update_qem_params(lr)
        out['states'].append(snap(prob))
    return out


def main():
    for name in models.QEM_CASES:
        for dtype in (t.float32, t.float64):
            out = run_case(name, dtype)
            tag = 'f32' if dtype == t.float32 else 'f64'
            path = os.path.join(HERE, f"{name}_{tag}.pt")
            t.save(out, path)
            last = out['states'][-1]
            print(f"{name:14s} {tag} Q params after {len(out['lrs'])} updates: "
                  f"{ {k: [round(float(x), 4) for x in v[0].reshape(-1)[:2]] for k, v in last['Q']['params'].items()} } "
                  f"-> {os.path.relpath(path, ROOT)} ({os.path.getsize(path)} B)")
    t.set_default_dtype(t.float32)


if __name__ == '__main__':
    main()
