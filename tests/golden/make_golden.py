"""Generate golden vectors by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py [case ...]  # writes tests/golden/<case>_<dtype>.pt (default: every case)

The reference sources are imported from /root/reference/src through oracle/refcompat.py
(runtime compat patches for torch 2.11, none on the logPQ arithmetic path; opt_einsum is
replaced by the greedy contraction-order shim).  This script cannot run on the
GPU box (no /root/reference there); its outputs are committed.

For each case of tests/models.py::CASES and each of float32 / float64 it stores
  inputs : sample / params+inputs / data tensors as (tensor, axes) pairs
  outputs: elbo (no_checkpoint; checkpoint and Split are asserted equal here, mirroring
           tests/test_problem_vs_itself.py:231-280), d elbo / d(sample, params), marginals
           (univariate + the case's joints), moments, and importance-sample indices drawn with
           EXPLICIT uniforms (torch.multinomial is swapped for the inverse-CDF rule of
           SURVEY.md Appendix A8 while the reference's own tree walk runs).
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
from alan.utils import generic_dims, generic_order, generic_getitem
from alan.Plate import flatten_tree

import models
from uniforms import UniformSource


def plain(x):
    """torchdim tensor -> (plain tensor, axes names)"""
    dims = generic_dims(x)
    return generic_order(x, dims).detach().clone(), tuple(str(d) for d in dims)


def named_plain(x):
    names = list(x.names)
    k = sum(n is not None for n in names)
    assert all(n is not None for n in names[:k])
    return x.detach().rename(None).clone(), tuple(names[:k])


def rebuild(tree, leaves, dimsof):
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out[k] = rebuild(v, leaves, dimsof)
        else:
            dims = generic_dims(v)
            leaf = generic_order(v, dims).detach().clone().requires_grad_()
            leaves[k] = leaf
            dimsof[k] = dims
            out[k] = generic_getitem(leaf, dims)
    return out


class PatchedMultinomial:
    """inverse-CDF categorical draw from explicit uniforms, shaped the way the reference's
    sample_Ks expects torch.multinomial's result (reduce_Ks.py:71-75)."""
    def __init__(self, source):
        self.source = source

    def __call__(self, logits, num_samples, replacement=True):
        dims = generic_dims(logits)
        Ndim = [d for d in dims if str(d) == 'N']
        pdims = {str(d): d for d in dims if str(d) != 'N'}
        canon = [pdims[n] for n in self.source.canon(pdims.keys())]
        u, _ = self.source.draw(pdims.keys())                       # [plates..., N]
        p = generic_order(logits, [*canon, *Ndim]).to(t.float64)    # [plates..., (N), C]
        if not Ndim:
            p = p.unsqueeze(-2)
        c = p.cumsum(-1)
        thr = (u * c[..., -1]).unsqueeze(-1)
        flat = (c < thr).sum(-1).clamp(max=p.shape[-1] - 1)         # [plates..., N]
        if Ndim:
            # value independent of the positional draw index; reference takes the diagonal
            out = flat.unsqueeze(0).expand(num_samples, *flat.shape)
            return generic_getitem(out, [slice(None), *canon, Ndim[0]])
        out = flat.movedim(-1, 0)
        return generic_getitem(out, [slice(None), *canon])


def run_case(name, dtype, seed=0):
    model, inputs_fn, kw, K, moms, joints, N = models.CASES[name]
    t.set_default_dtype(dtype)
    t.manual_seed(seed)
    inp = inputs_fn(**kw, seed=seed, dtype=dtype)
    P, Q = model(alan)
    bp = alan.BoundPlate(P, inp['platesizes'], inputs=inp['inputs'])
    bq = alan.BoundPlate(Q, inp['platesizes'], inputs=inp['inputs'],
                         extra_opt_params={k: v.clone() for k, v in inp['params'].items()})
    prob = alan.Problem(bp, bq, inp['data'])
    tree, g2K = prob.Q._sample(K, False, alan.PermutationSampler, prob.all_platedims)
    leaves, dimsof = {}, {}
    tree = rebuild(tree, leaves, dimsof)
    s = alan.Sample(prob, tree, g2K, alan.PermutationSampler, reparam=True)

    elbo = s.elbo_vi(computation_strategy=alan.no_checkpoint)
    params = dict(prob.Q._opt_params.to_dict())
    pnames = list(params.keys())
    lnames = list(leaves.keys())
    grads = t.autograd.grad(elbo, [leaves[k] for k in lnames] + [params[k] for k in pnames], allow_unused=True)

    # computation strategies agree (tests/test_problem_vs_itself.py:231-280)
    e_ck = s.elbo_vi(computation_strategy=alan.checkpoint)
    assert t.isclose(elbo, e_ck, rtol=1e-5 if dtype == t.float32 else 1e-10), (elbo, e_ck)

    out = {
        'case': name, 'K': K, 'dtype': str(dtype), 'platesizes': inp['platesizes'],
        'sample': {k: (leaves[k].detach().clone(), tuple(str(d) for d in dimsof[k])) for k in lnames},
        'params': {k: named_plain(v) for k, v in inp['params'].items()},
        'inputs': {k: named_plain(v) for k, v in inp['inputs'].items()},
        'data': {k: named_plain(v) for k, v in inp['data'].items()},
        'elbo': elbo.detach().clone(),
        'grad_sample': {k: g.detach().clone() for k, g in zip(lnames, grads[:len(lnames)]) if g is not None},
        'grad_params': {k: g.detach().rename(None).clone() for k, g in zip(pnames, grads[len(lnames):])
                        if g is not None},
    }

    marg = s.marginals(joints=joints, computation_strategy=alan.no_checkpoint)
    out['marginals'] = {tuple(sorted(k)): plain(v) for k, v in marg.weights.items()}

    mlist = [((v,), alan.moments.RawMoment(models.MOMENT_FUNCS[f])) for v, f in moms]
    mres = s._moments_uniform_input(mlist)
    out['moments'] = [plain(m) for m in mres]
    out['moment_specs'] = moms

    if N is not None:
        src = UniformSource(seed + 1, N, inp['platesizes'], list(inp['platesizes']))
        real = t.multinomial
        rk = sys.modules['alan.reduce_Ks']
        # make joint-K raveling order deterministic: the reference iterates a Python set of Dims
        # (reduce_Ks.py:276); sort each step's Ks by the order they were asked to be summed.
        orig_collect = rk.collect_lps

        def collect_sorted(lps, Ks_to_sum):
            result, all_reduced, Ks_to_sample = orig_collect(lps, Ks_to_sum)
            pos = {id(k): i for i, k in enumerate(Ks_to_sum)}
            Ks_to_sample = [tuple(sorted(ks, key=lambda k: pos[id(k)])) for ks in Ks_to_sample]
            return result, all_reduced, Ks_to_sample
        rk.collect_lps = collect_sorted
        t.multinomial = PatchedMultinomial(src)
        try:
            idxs, N_dim = s._importance_sample_idxs(N, alan.no_checkpoint)
        finally:
            t.multinomial = real
            rk.collect_lps = orig_collect
        out['N'] = N
        out['uniform_seed'] = seed + 1
        out['indices'] = {g: plain(v) for g, v in idxs.items()}
    return out


def main():
    only = sys.argv[1:]                       # optional: the cases to (re)generate; default all
    for name in models.CASES:
        if only and name not in only:
            continue
        for dtype in (t.float32, t.float64):
            out = run_case(name, dtype)
            tag = 'f32' if dtype == t.float32 else 'f64'
            path = os.path.join(HERE, f"{name}_{tag}.pt")
            t.save(out, path)
            print(f"{name:18s} {tag}  elbo={out['elbo'].item():+.10f}  -> {os.path.relpath(path, ROOT)} "
                  f"({os.path.getsize(path)} B)")
    t.set_default_dtype(t.float32)


if __name__ == '__main__':
    main()
