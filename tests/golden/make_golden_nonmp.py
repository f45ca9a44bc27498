"""Golden vectors of the reference's global importance sampling (`SampleNonMP`), from the
UNMODIFIED reference in the build container.

    python tests/golden/make_golden_nonmp.py            # writes tests/golden/nonmp_<case>_<dtype>.pt

Per case (tests/models.py::CASES without the Timeseries one) and dtype: K independent joint draws from Q
(`IndependentSampler`), then through `alan.SampleNonMP.SampleNonMP`:
  lpq [K] (`logpq`), elbo (`elbo_vi`), d elbo / d(sample, params), moments (`_moments_uniform_input`).
The categorical draw of `importance_sample` goes through torch.multinomial's RNG stream upstream; its parity is
defined on lpq (the weights) plus the explicit-uniform inverse-CDF rule of SURVEY.md Appendix A8.
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
from alan.Sampler import IndependentSampler
from alan.SampleNonMP import SampleNonMP

import models
from make_golden import plain, named_plain, rebuild

NONMP_CASES = ['cfg1_lgl', 'cfg1_lglp', 'cfg2_movielens', 'cfg3_radon', 'model1', 'ref_bernoulli']


def run_case(name, dtype, seed=0):
    model, inputs_fn, kw, K, moms, joints, N = models.CASES[name]
    K = max(K, 8)
    t.set_default_dtype(dtype)
    t.manual_seed(seed + 11)
    inp = inputs_fn(**kw, seed=seed, dtype=dtype)
    P, Q = model(alan)
    bp = alan.BoundPlate(P, inp['platesizes'], inputs=inp['inputs'])
    bq = alan.BoundPlate(Q, inp['platesizes'], inputs=inp['inputs'],
                         extra_opt_params={k: v.clone() for k, v in inp['params'].items()})
    prob = alan.Problem(bp, bq, inp['data'])
    tree, g2K = prob.Q._sample(K, False, IndependentSampler, prob.all_platedims)
    leaves, dimsof = {}, {}
    tree = rebuild(tree, leaves, dimsof)
    s = SampleNonMP(prob, tree, g2K, reparam=True)
    elbo = s.elbo_vi()
    params = dict(prob.Q._opt_params.to_dict())
    pnames, lnames = list(params.keys()), list(leaves.keys())
    grads = t.autograd.grad(elbo, [leaves[k] for k in lnames] + [params[k] for k in pnames], allow_unused=True)
    lpq = s.logpq(s.detached_sample).order(s.Kdim).detach().clone()
    out = {
        'case': name, 'K': K, 'dtype': str(dtype), 'platesizes': inp['platesizes'],
        'sample': {k: (leaves[k].detach().clone(), tuple(str(d) for d in dimsof[k])) for k in lnames},
        'params': {k: named_plain(v) for k, v in inp['params'].items()},
        'inputs': {k: named_plain(v) for k, v in inp['inputs'].items()},
        'data': {k: named_plain(v) for k, v in inp['data'].items()},
        'lpq': lpq, 'elbo': elbo.detach().clone(),
        'grad_sample': {k: g.detach().clone() for k, g in zip(lnames, grads[:len(lnames)]) if g is not None},
        'grad_params': {k: g.detach().rename(None).clone() for k, g in zip(pnames, grads[len(lnames):]) if g is not None},
    }
    mlist = [((v,), alan.moments.RawMoment(models.MOMENT_FUNCS[f])) for v, f in moms]
    out['moments'] = [plain(m) for m in s._moments_uniform_input(mlist)]
    out['moment_specs'] = moms
    return out


def main():
    for name in NONMP_CASES:
        for dtype in (t.float32, t.float64):
            out = run_case(name, dtype)
            tag = 'f32' if dtype == t.float32 else 'f64'
            path = os.path.join(HERE, f"nonmp_{name}_{tag}.pt")
            t.save(out, path)
            print(f"{name:18s} {tag}  elbo={out['elbo'].item():+.10f}  -> {os.path.relpath(path, ROOT)} "
                  f"({os.path.getsize(path)} B)")
    t.set_default_dtype(t.float32)


if __name__ == '__main__':
    main()
