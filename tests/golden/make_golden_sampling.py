"""Golden samples of Q (row f-1) from the UNMODIFIED reference walk, with explicit base noise.

    python tests/golden/make_golden_sampling.py        # writes tests/golden/qsample_<case>_<dtype>.pt

`BoundPlate._sample` -> `Plate.sample` -> `sample_gdt` -> `Sampler.resample_scope` / `Timeseries.sample`
(/root/reference/src/alan/BoundPlate.py:338-363, Plate.py:93-143, dist.py:23-72, Sampler.py:85-116,
Timeseries.py:89-123) run as they are; only the two primitives that consume the RNG are replaced by explicit-noise
versions keyed by name (the reference's RNG stream depends on the iteration order of Python sets of Dims):
  * `PermutationSampler.perm(dims, Kdim)` -> argsort of the supplied float64 uniforms (Sampler.py:143-148 draws
    Uniform(0,1) and argsorts: same rule, explicit uniforms);
  * `Dist.sample` -> the closed-form transform of supplied base noise (Normal: loc + scale * eps, what
    torch.distributions' rsample computes), with the arguments resolved by the reference's own `paramname2val`.
Cases cover: mixture Q with a lambda on a parent (cfg1), plated + global latents (cfg2), a Group and three plate
levels (cfg3), a Group member depending on an earlier member and a latent scale (model1), a correlated Q (ref_corr_q)
and a Timeseries drawn through the reference's T-step loop (the prior P of cfg4).
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
from oracle.sample_oracle import transform, perm_from_uniform

CASES = {
    # name: (models case, which side is sampled, K)
    'cfg1_lgl': ('cfg1_lgl', 'Q', 5), 'cfg1_lglp': ('cfg1_lglp', 'Q', 4), 'cfg2_movielens': ('cfg2_movielens', 'Q', 6),
    'cfg3_radon': ('cfg3_radon', 'Q', 4), 'model1': ('model1', 'Q', 5), 'ref_corr_q': ('ref_corr_q', 'Q', 6),
    'cfg4_timeseries_P': ('cfg4_timeseries', 'P', 5),
    # IndependentSampler (Sampler.py:162-169, the sampler of Problem.sample_nonmp): child particle k sees parent particle k
    'cfg1_lglp_indep': ('cfg1_lglp', 'Q', 4, 'indep'), 'cfg3_radon_indep': ('cfg3_radon', 'Q', 4, 'indep'),
}


def named_plain(x):
    names = list(x.names)
    k = sum(n is not None for n in names)
    return x.detach().rename(None).clone(), tuple(names[:k])


class Noise:
    """named base noise, drawn lazily with a fixed seed per key"""
    def __init__(self, seed, dtype):
        self.seed, self.dtype, self.store, self.kinds = seed, dtype, {}, {}

    def get(self, key, shape, kind):
        if key not in self.store:
            g = t.Generator().manual_seed(self.seed * 1009 + len(self.store))
            if kind == 'normal':
                x = t.randn(shape, dtype=t.float64, generator=g).to(self.dtype)
            elif kind == 'perm':
                x = t.rand(shape, dtype=t.float64, generator=g)
            else:
                x = t.rand(shape, dtype=t.float64, generator=g).to(self.dtype)
            self.store[key] = x
            self.kinds[key] = kind
        assert tuple(self.store[key].shape) == tuple(shape), (key, shape, self.store[key].shape)
        return self.store[key]


def run_case(name, dtype, seed=0):
    case, side, K = CASES[name][:3]
    indep = len(CASES[name]) > 3 and CASES[name][3] == 'indep'
    model, inputs_fn, kw, _, _, _, _ = models.CASES[case]
    t.set_default_dtype(dtype)
    t.manual_seed(seed)
    inp = inputs_fn(**kw, seed=seed, dtype=dtype)
    P, Q = model(alan)
    plate = Q if side == 'Q' else P
    extra = {k: v.clone() for k, v in inp['params'].items()} if side == 'Q' else {}
    bp = alan.BoundPlate(plate, inp['platesizes'], inputs=inp['inputs'], extra_opt_params=extra)
    all_platedims = bp.all_platedims if hasattr(bp, 'all_platedims') else None
    if all_platedims is None:
        from functorch.dim import Dim
        all_platedims = {k: Dim(k, v) for k, v in inp['platesizes'].items()}
    plate_order = list(inp['platesizes'])
    noise = Noise(seed + 17, dtype)
    from oracle.sample_oracle import transform as tf
    from alan_b200.sampling import NOISE_KIND

    # ---- which Dist object belongs to which variable
    dist2var, ts_vars, ts_count = {}, set(), {}

    def walk(pl):
        for k, v in pl.flat_prog.items():
            if isinstance(v, alan.Plate):
                walk(v)
            elif getattr(v, 'is_timeseries', False):
                dist2var[id(v.trans)] = k
                ts_vars.add(k)
            elif hasattr(v, 'dist'):
                dist2var[id(v)] = k
    walk(bp.plate)
    cur = {'group': None}
    S = sys.modules['alan.Sampler']
    D = sys.modules['alan.dist']
    orig_rs = S.Sampler.resample_scope.__func__
    orig_perm = S.PermutationSampler.perm
    orig_sample = D.Dist.sample

    def resample_scope(cls, scope, active_platedims, Kdim):
        cur['group'] = str(Kdim)[2:]
        return orig_rs(cls, scope, active_platedims, Kdim)

    def perm(dims, Kdim):
        plates = [d for n in plate_order for d in dims if str(d) == n]
        own = str(Kdim) == 'K_' + str(cur['group'])
        key = (cur['group'], 'timeseries' if own else str(Kdim))
        u = noise.get(key, [d.size for d in plates] + [Kdim.size], 'perm')
        p = perm_from_uniform(u, 0).movedim(-1, 0)                         # [K, plates...]
        return generic_getitem(p, [slice(None), *plates])

    def sample(self, scope, reparam, active_platedims, K_dim, timeseries_perm=None):
        var = dist2var[id(self)]
        args = self.paramname2val(scope)
        family = self.dist.__name__
        if var in ts_vars:
            # called once per time step with the plates ABOVE the time axis (Timeseries.py:113)
            tix = ts_count.get(var, 0)
            ts_count[var] = tix + 1
            T_name = [n for n in plate_order if n not in [str(d) for d in active_platedims]]
            full_plates = [d for n in plate_order for d in all_platedims.values() if str(d) == n and
                           (d in active_platedims or n == ts_plate[var])]
            ev = event_shape(args)
            e = noise.get(var, [d.size for d in full_plates] + [K_dim.size] + ev, NOISE_KIND[family])
            tpos = [str(d) for d in full_plates].index(ts_plate[var])
            e = e.select(tpos, tix)
            plates = [d for d in full_plates if str(d) != ts_plate[var]]
        else:
            plates = [d for n in plate_order for d in active_platedims if str(d) == n]
            ev = event_shape(args)
            e = noise.get(var, [d.size for d in plates] + [K_dim.size] + ev, NOISE_KIND[family])
        e_td = generic_getitem(e, [*plates, K_dim, *([slice(None)] * len(ev))])
        return tf(family, args, e_td)

    def event_shape(args):
        shp = ()
        for v in args.values():
            if hasattr(v, 'shape'):                                   # torch or functorch.dim tensor: positional shape
                shp = t.broadcast_shapes(shp, tuple(v.shape))
        return list(shp)

    # the plate that carries each Timeseries (its time axis)
    ts_plate = {}

    def find_ts(pl, name):
        for k, v in pl.flat_prog.items():
            if isinstance(v, alan.Plate):
                find_ts(v, k)
            elif getattr(v, 'is_timeseries', False):
                ts_plate[k] = name
    find_ts(bp.plate, None)

    S.Sampler.resample_scope = classmethod(resample_scope)
    S.PermutationSampler.perm = staticmethod(perm)
    D.Dist.sample = sample
    try:
        tree, g2K = bp._sample(K, False, S.IndependentSampler if indep else alan.PermutationSampler, all_platedims)
    finally:
        S.Sampler.resample_scope = classmethod(orig_rs)
        S.PermutationSampler.perm = orig_perm
        D.Dist.sample = orig_sample
    flat = flatten_tree(tree)
    out_samples = {}
    for k, v in flat.items():
        dims = generic_dims(v)
        pl = [d for n in plate_order for d in dims if str(d) == n]
        kd = [d for d in dims if str(d).startswith('K_')]
        assert len(kd) == 1
        out_samples[k] = (generic_order(v, [*pl, *kd]).detach().clone(), tuple(str(d) for d in [*pl, *kd]))
    return {
        'case': case, 'side': side, 'K': K, 'dtype': str(dtype), 'platesizes': inp['platesizes'],
        'sampler_mode': 2 if indep else 0,
        'params': {k: named_plain(v) for k, v in (inp['params'].items() if side == 'Q' else [])},
        'inputs': {k: named_plain(v) for k, v in inp['inputs'].items()},
        'noise': dict(noise.store), 'noise_kinds': dict(noise.kinds),
        'samples': out_samples,
    }


def main():
    for name in (sys.argv[1:] or CASES):
        for dtype in (t.float32, t.float64):
            out = run_case(name, dtype)
            tag = 'f32' if dtype == t.float32 else 'f64'
            path = os.path.join(HERE, f"qsample_{name}_{tag}.pt")
            t.save(out, path)
            print(f"{name:20s} {tag} vars={list(out['samples'])} noise={len(out['noise'])} -> {os.path.relpath(path, ROOT)} "
                  f"({os.path.getsize(path)} B)")
    t.set_default_dtype(t.float32)


if __name__ == '__main__':
    main()
