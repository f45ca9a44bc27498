"""Golden outputs of prediction (row f-3) from the UNMODIFIED reference, with explicit base noise.

    python tests/golden/make_golden_predict.py        # writes tests/golden/predict_<case>_<dtype>.pt

`ImportanceSample.extend` -> `Plate.sample_extended` -> `Dist.sample_extended` and
`ExtendedImportanceSample.predictive_ll` (/root/reference/src/alan/ImportanceSample.py:43-177, Plate.py:145-215,
dist.py:234-294) run as they are on a posterior sample built by the reference's own `index_into_sample`; only
`TorchDimDist.sample`, the one primitive that consumes the RNG, is replaced by the closed-form transform of supplied
base noise keyed by variable name (Normal: loc + scale * eps; Bernoulli: u < p).
Cases: MovieLens-shaped (two nested plates extended, a feature input, Bernoulli data through a matrix-vector lambda)
and the radon-shaped three-level hierarchy (the innermost plate extended, two covariate inputs, Normal data).
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
from functorch.dim import Dim

import models
from oracle.sample_oracle import transform
from alan_b200.sampling import NOISE_KIND

# name: (models case, K, N, extended plate sizes, builder kwargs)
CASES = {
    'cfg2_movielens': ('cfg2_movielens', 5, 7, {'plate_1': 15, 'plate_2': 8}),
    'cfg3_radon': ('cfg3_radon', 4, 6, {'Zips': 8, 'Counties': 6}),
}


def named_plain(x):
    names = list(x.names)
    k = sum(n is not None for n in names)
    return x.detach().rename(None).clone(), tuple(names[:k])


def extend_named(x, ext_sizes, gen, kind, dtype):
    """a named tensor grown to the extended plate sizes: the original block kept, the rest fresh"""
    names = list(x.names)
    shape = [ext_sizes.get(n, s) if n is not None else s for n, s in zip(names, x.shape)]
    if kind == 'binary':
        y = (t.rand(shape, generator=gen) < 0.4).to(dtype)
    else:
        y = t.randn(shape, generator=gen, dtype=t.float64).to(dtype)
    y[tuple(slice(0, s) for s in x.shape)] = x.rename(None)
    return y.refine_names(*names)


def run_case(name, dtype, seed=0):
    case, K, N, ext = CASES[name]
    model, inputs_fn, kw, _, _, _, _ = models.CASES[case]
    t.set_default_dtype(dtype)
    t.manual_seed(seed)
    inp = inputs_fn(**kw, seed=seed, dtype=dtype)
    P, Q = model(alan)
    bP = alan.BoundPlate(P, inp['platesizes'], inputs=inp['inputs'])
    bQ = alan.BoundPlate(Q, inp['platesizes'], inputs=inp['inputs'], extra_opt_params={k: v.clone() for k, v in inp['params'].items()})
    prob = alan.Problem(bP, bQ, inp['data'])
    s = prob.sample(K)
    isamp = s.importance_sample(N)
    ext_sizes = {**inp['platesizes'], **ext}
    g = t.Generator().manual_seed(seed + 5)
    binary = lambda x: bool(((x.rename(None) == 0) | (x.rename(None) == 1)).all())
    ext_inputs = {k: extend_named(v, ext_sizes, g, 'binary' if binary(v) else 'real', dtype) for k, v in inp['inputs'].items()}
    ext_data = {k: extend_named(v, ext_sizes, g, 'binary' if binary(v) else 'real', dtype) for k, v in inp['data'].items()}

    # ---- explicit noise in place of TorchDimDist.sample
    D = sys.modules['alan.dist']
    TDD = sys.modules['alan.TorchDimDist']
    noise, kinds, cur = {}, {}, {}
    dist2var = {}

    def walk(pl):
        for k, v in pl.flat_prog.items():
            if isinstance(v, alan.Plate):
                walk(v)
            else:
                dist2var[id(v)] = k
    walk(bP.plate)
    plate_order = list(ext_sizes)
    orig_ext = D.Dist.sample_extended
    orig_sample = TDD.TorchDimDist.sample

    def sample_extended(self, *a, **kw):
        cur['var'], cur['family'] = dist2var[id(self)], self.dist.__name__
        return orig_ext(self, *a, **kw)

    def tdd_sample(self, reparam, sample_dims, sample_shape):
        var, family = cur['var'], cur['family']
        plates = [d for n in plate_order for d in sample_dims if str(d) == n]
        Nd = [d for d in sample_dims if str(d) == 'N']
        args = dict(self.kwargs_torchdim)
        ev = ()
        for v in args.values():
            if hasattr(v, 'shape'):
                ev = t.broadcast_shapes(ev, tuple(v.shape))
        kind = NOISE_KIND[family]
        shape = [d.size for d in plates] + [Nd[0].size] + list(ev)
        gg = t.Generator().manual_seed(1009 * (seed + 3) + len(noise))
        e = (t.randn(shape, generator=gg, dtype=t.float64) if kind == 'normal' else t.rand(shape, generator=gg, dtype=t.float64)).to(dtype)
        noise[var], kinds[var] = e, kind
        e_td = generic_getitem(e, [*plates, Nd[0], *([slice(None)] * len(ev))])
        return transform(family, args, e_td)

    D.Dist.sample_extended = sample_extended
    TDD.TorchDimDist.sample = tdd_sample
    try:
        ext_s = isamp.extend(dict(ext_sizes), ext_inputs)
    finally:
        D.Dist.sample_extended = orig_ext
        TDD.TorchDimDist.sample = orig_sample
    pll = ext_s.predictive_ll(dict(ext_data))

    def dump(flat, Ndim):
        out = {}
        for k, v in flat.items():
            dims = generic_dims(v)
            pl = [d for n in plate_order for d in dims if str(d) == n]
            nd = [d for d in dims if str(d) == 'N']
            out[k] = (generic_order(v, [*pl, *nd]).detach().clone(), tuple(str(d) for d in [*pl, *nd]))
        return out
    return {
        'case': case, 'K': K, 'N': N, 'dtype': str(dtype), 'platesizes': inp['platesizes'], 'ext_sizes': ext_sizes,
        'data': {k: named_plain(v) for k, v in inp['data'].items()},
        'ext_inputs': {k: named_plain(v) for k, v in ext_inputs.items()},
        'ext_data': {k: named_plain(v) for k, v in ext_data.items()},
        'post': dump(isamp.samples_flatdict, isamp.Ndim),
        'noise': noise, 'noise_kinds': kinds,
        'extended': dump(ext_s.samples_flatdict, ext_s.Ndim),
        'pll': {k: v.detach().clone() for k, v in pll.items()},
    }


def main():
    for name in CASES:
        for dtype in (t.float32, t.float64):
            out = run_case(name, dtype)
            tag = 'f32' if dtype == t.float32 else 'f64'
            path = os.path.join(HERE, f"predict_{name}_{tag}.pt")
            t.save(out, path)
            print(f"{name:16s} {tag} extended={list(out['extended'])} pll={ {k: float(v) for k, v in out['pll'].items()} } "
                  f"-> {os.path.relpath(path, ROOT)} ({os.path.getsize(path)} B)")
    t.set_default_dtype(t.float32)


if __name__ == '__main__':
    main()
