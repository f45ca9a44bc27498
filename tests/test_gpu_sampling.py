"""Ancestral sampling of Q on the GPU (alan_b200/sampling.py, csrc/sampling.cuh) through the
C ABI, against
 (1) the golden samples of the UNMODIFIED reference walk with explicit base noise (tests/golden/qsample_*.pt):
     permutations bit-exact, draws within one rounding of `loc + scale * eps` (the device contracts it to an FMA);
 (2) the oracle (oracle/sample_oracle.py) at sizes the goldens do not reach: CategoricalSampler, BASELINE-shaped
     MovieLens (300 x 5, K = 30), a Timeseries of T = 1000 steps drawn in one launch;
 (3) `Problem.sample(K)` end to end: the sample feeds `elbo_rws` / `marginals` like a reference sample."""
import os

import pytest
import torch as t

import models
from alan_b200 import model as M
from alan_b200.named import NT
from golden_io import GOLDEN_DIR, TAGS

pytestmark = pytest.mark.gpu
CASES = ['cfg1_lgl', 'cfg1_lglp', 'cfg2_movielens', 'cfg3_radon', 'model1', 'ref_corr_q', 'cfg4_timeseries_P',
         'cfg1_lglp_indep', 'cfg3_radon_indep']          # *_indep: IndependentSampler (Problem.sample_nonmp)


def load(case, tag):
    return t.load(os.path.join(GOLDEN_DIR, f"qsample_{case}_{tag}.pt"), weights_only=False)


def plate_of(g):
    P, Q = models.build(g['case'], M, t.float64 if 'float64' in g['dtype'] else t.float32)
    return Q if g['side'] == 'Q' else P


def params_of(g):
    return {k: NT(v[0], v[1]) for d in (g['inputs'], g['params']) for k, v in d.items()}


def close(a, b, tag):
    tol = 2e-6 if tag == 'f32' else 1e-13
    return float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_sampling_matches_reference_walk(case, tag):
    from alan_b200.sampling import QSampler, PermutationSampler, IndependentSampler
    g = load(case, tag)
    ip = params_of(g)
    sampler = IndependentSampler if g.get('sampler_mode', 0) == 2 else PermutationSampler
    qs = QSampler(plate_of(g), ip, g['platesizes'], g['K'], sampler, TAGS[tag], 'cuda:0')
    out = qs.run(ip, noise={k: g['noise'][k] for k in qs.noise_shapes()})
    assert set(out) == set(g['samples'])
    for var, (ref, axes) in g['samples'].items():
        assert close(out[var].order(axes).t.cpu(), ref, tag), var


@pytest.mark.parametrize("mode", ['permutation', 'categorical'])
@pytest.mark.parametrize("side", ['P', 'Q'])
@pytest.mark.parametrize("tag", list(TAGS))
def test_sampling_movielens_full_size_vs_oracle(tag, side, mode):
    """cfg-2 shape (300 users x 5 films, d = 18, K = 30) under both samplers.  The prior side draws z from the
    permuted / resampled mu_z and psi_z particles and obs from the permuted z particles (a matrix-vector lambda)."""
    from alan_b200.sampling import QSampler, PermutationSampler, CategoricalSampler
    from oracle.sample_oracle import sample_q
    dt = TAGS[tag]
    inp = models.movielens_inputs(dtype=dt)
    P, Q = models.build('cfg2_movielens', M, dt)
    from alan_b200.named import from_torch_named
    if side == 'Q':
        plate, ip = Q, {k: from_torch_named(v) if any(v.names) else NT(v, ()) for k, v in inp['params'].items()}
    else:
        plate, ip = P, {k: from_torch_named(v) for k, v in inp['inputs'].items()}
    S = PermutationSampler if mode == 'permutation' else CategoricalSampler
    K = 30
    qs = QSampler(plate, ip, inp['platesizes'], K, S, dt, 'cuda:0')
    noise = qs.make_noise('cuda:0', seed=11)
    out = qs.run(ip, noise=noise)
    ref = sample_q(plate, ip, {k: v.cpu() for k, v in noise.items()}, K, S.mode, dt)
    assert set(out) == set(ref)
    for var, r in ref.items():
        got = out[var].order(r.axes).t.cpu()
        if var == 'obs':                                   # a Bernoulli draw: u < sigmoid(z . x), flips only at ties
            assert float((got != r.t).double().mean()) < 1e-4
        else:
            assert close(got, r.t, tag), var


@pytest.mark.parametrize("tag", list(TAGS))
def test_sampling_timeseries_T1000_vs_oracle(tag):
    """cfg-4 prior (T = 1000, K = 16): the reference's T-step Python loop as one launch."""
    from alan_b200.sampling import QSampler, PermutationSampler
    from oracle.sample_oracle import sample_q
    dt = TAGS[tag]
    P, _ = models.build('cfg4_timeseries', M, dt)
    qs = QSampler(P, {}, {'T': 1000}, 16, PermutationSampler, dt, 'cuda:0')
    assert sum(type(op).__name__ == 'TsSampleOp' for op in qs.plan.programs[0]) == 1
    noise = qs.make_noise('cuda:0', seed=5)
    out = qs.run({}, noise=noise)
    ref = sample_q(P, {}, {k: v.cpu() for k, v in noise.items()}, 16, 0, dt)
    # 1000 dependent steps: the one-rounding difference of every step is carried along the chain (|0.9| < 1 damps it)
    tol = 2e-5 if tag == 'f32' else 1e-12
    for var, r in ref.items():
        got = out[var].order(r.axes).t.cpu()
        assert float((got - r.t).abs().max()) <= tol * max(1.0, float(r.t.abs().max())), var


def test_problem_sample_feeds_the_logpq_path():
    """Problem.sample(K) -> elbo_rws().backward() / marginals() / importance_sample(): moments of the draw match the
    Q parameters, K axes lead, the same seed gives the same sample, and the log-evidence equals the oracle's on the
    drawn sample."""
    from alan_b200.problem import Problem
    from oracle import logpq_oracle as O
    from golden_io import load as load_lp, rel_err
    g = load_lp("cfg2_movielens", "f32")
    P, Q = models.build("cfg2_movielens", M, t.float32)
    nt = lambda d, rg=False: {k: NT(v[0].clone().requires_grad_(rg), v[1]) for k, v in d.items()}
    params = nt(g["params"], True)
    prob = Problem(P, Q, nt(g["data"]), inputs=nt(g["inputs"]), params=params, device="cuda:0")
    K = 30
    s = prob.sample(K, reparam=False, seed=1)
    s2 = prob.sample(K, reparam=False, seed=1)
    s3 = prob.sample(K, reparam=False, seed=2)
    for k, v in s.sample.items():
        assert v.axes[0].startswith('K_') and v.t.shape[0] == K and v.t.is_cuda
        assert t.equal(v.t, s2.sample[k].t) and not t.equal(v.t, s3.sample[k].t)
    z = s.sample['z']
    zl = params['z_loc'].t.detach().cuda()
    zs = params['z_ls'].t.detach().cuda().exp()
    zz = (z.t - zl) / zs                                                   # standard normal under Q
    assert abs(float(zz.mean())) < 0.03 and abs(float(zz.std()) - 1) < 0.03
    L = s.elbo_rws()
    host = {k: NT(v.t.detach().cpu(), v.axes) for k, v in s.sample.items()}
    ip = {k: NT(v.t.detach(), v.axes) for k, v in prob.inputs_params().items()}
    ref = O.elbo(P, Q, host, ip, prob.data)
    assert rel_err(L.detach().cpu(), ref) < 1e-5
    L.backward()
    assert all(p.t.grad is not None and bool(t.isfinite(p.t.grad).all()) for p in params.values())
    m = s.marginals()
    for key, w in m.weights.items():
        kd = tuple(i for i, a in enumerate(w.axes) if a.startswith('K_'))
        assert float((w.t.sum(kd) - 1).abs().max()) < 1e-4, key
    post = s.importance_sample(7, seed=0)
    assert post['z'].t.shape[0] == 7


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", ['mvn', 'mvn2'])
def test_mvn_sampling_vs_oracle(case, tag):
    """MultivariateNormal draws (row f-2): loc + scale_tril eps with the factor computed on the device (csrc/mvn.cuh)
    from covariance / precision / scale_tril arguments, against the oracle on the same base noise."""
    from test_sampling_cpu import mvn_sampling_case
    qs, ip, noise, want = mvn_sampling_case(case, tag)
    qs.device = 'cuda:0'
    out = qs.run(ip, noise=noise)
    tol = 3e-6 if tag == 'f32' else 1e-12
    for var in want:
        mine, ref = out[var].order(want[var].axes).t.cpu(), want[var].t
        assert (mine - ref).abs().max() <= tol * max(1.0, float(ref.abs().max())), var
