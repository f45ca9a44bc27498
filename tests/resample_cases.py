"""BASELINE-size resampling cases shared by the golden generator (tests/golden/make_golden_resampling.py) and the GPU
test: the inputs and the K-sample are REGENERATED from seeds on both sides (CPU torch generators + the sampling
oracle), so the fixtures hold nothing but the reference's resampled indices."""
import torch as t

import models
from alan_b200 import model as M
from alan_b200.named import NT, from_torch_named

# name: (models case, inputs kwargs, K, N)          cfg-2 / cfg-3 shapes of BASELINE.json, importance_sample(N=100)
CASES = {
    'cfg2_movielens_300x5_K30': ('cfg2_movielens', dict(M=300, N=5, d=18), 30, 100),
    'cfg3_radon_12x16x10_K10': ('cfg3_radon', dict(S=12, C=16, Z=10), 10, 100),
}


def build(name, dtype, seed=0):
    """(P, Q, inp, sample {var: NT[plates..., K, *event]}) -- the sample is a draw from Q through the sampling oracle
    with seeded base noise (deterministic on CPU)."""
    from oracle.sample_oracle import sample_q
    from alan_b200.sampling import QSampler, PermutationSampler
    case, kw, K, N = CASES[name]
    old = t.get_default_dtype()
    t.set_default_dtype(dtype)            # the generator ran under the case's dtype: seeded `rand` streams depend on it
    try:
        return _build(case, kw, K, N, dtype, seed)
    finally:
        t.set_default_dtype(old)


def _build(case, kw, K, N, dtype, seed):
    from oracle.sample_oracle import sample_q
    from alan_b200.sampling import QSampler, PermutationSampler
    inp = models.CASES[case][1](**kw, seed=seed, dtype=dtype)
    P, Q = models.build(case, M, dtype)
    nt = lambda d: {k: from_torch_named(v) if any(n is not None for n in v.names) else NT(v, ()) for k, v in d.items()}
    ip = {**nt(inp['inputs']), **nt(inp['params'])}
    qs = QSampler(Q, ip, inp['platesizes'], K, PermutationSampler, dtype)
    g = t.Generator().manual_seed(1000 + seed)
    noise = {}
    for key, (kind, shape, dt) in qs.noise_shapes().items():
        x = (t.randn if kind == 'normal' else t.rand)(shape, dtype=t.float64, generator=g)
        noise[key] = x if kind == 'perm' else x.to(dtype)
    sample = sample_q(Q, ip, noise, K, 0, dtype)
    return P, Q, inp, sample, K, N
