"""Prediction on the GPU (alan_b200/predict.py) through the C ABI, against the goldens of the
UNMODIFIED reference (`ImportanceSample.extend`, `ExtendedImportanceSample.predictive_ll` with explicit base noise,
tests/golden/make_golden_predict.py) and, at BASELINE cfg-2 size, against the oracle (oracle/predict_oracle.py)."""
import os

import pytest
import torch as t

import models
from alan_b200 import model as M
from alan_b200.named import NT, from_torch_named
from golden_io import GOLDEN_DIR, TAGS

pytestmark = pytest.mark.gpu
CASES = ['cfg2_movielens', 'cfg3_radon']


def load(case, tag):
    return t.load(os.path.join(GOLDEN_DIR, f"predict_{case}_{tag}.pt"), weights_only=False)


def nts(d):
    return {k: NT(v[0], v[1]) for k, v in d.items()}


def close(a, b, tag, scale=1.0):
    tol = (3e-6 if tag == 'f32' else 1e-12) * scale
    return float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_extend_and_predictive_ll_vs_reference(case, tag):
    from alan_b200.predict import Extender, PredictiveLL
    dt = TAGS[tag]
    g = load(case, tag)
    P, _ = models.build(g['case'], M, dt)
    post, data, ext_in, ext_data = nts(g['post']), nts(g['data']), nts(g['ext_inputs']), nts(g['ext_data'])
    ex = Extender(P, post, data, g['ext_sizes'], ext_in, g['N'], dt, 'cuda:0')
    ext = ex.run(post, data, ext_in, noise=g['noise'])
    assert set(ext) == set(g['extended'])
    for k, (ref, axes) in g['extended'].items():
        got = ext[k].order(axes).t.cpu()
        if g['noise_kinds'][k] == 'uniform':                     # Bernoulli draws: u < p, flips only at ties
            assert float((got != ref).double().mean()) < 1e-3, k
        else:
            assert close(got, ref, tag), k
    latents = {k: NT(*g['extended'][k]) for k in g['extended'] if k not in data}
    pl = PredictiveLL(P, latents, ext_data, g['platesizes'], ext_in, g['N'], dt, 'cuda:0')
    out = pl.run(latents, ext_data, ext_in)
    for k, ref in g['pll'].items():
        assert close(out[k].cpu(), ref, tag, scale=3.0), k


def test_prediction_through_the_api_at_cfg2_size():
    """Problem.sample(K).importance_sample(N).extend(...).predictive_ll(...) on the MovieLens shape (300 x 5 extended
    to 360 x 8): the original block survives the extension bit for bit, and the predictive log-likelihood equals the
    oracle's on the same extended sample."""
    from alan_b200.problem import Problem
    from oracle import predict_oracle as PO
    dt = t.float32
    inp = models.movielens_inputs(dtype=dt)
    P, Q = models.build('cfg2_movielens', M, dt)
    nt = lambda d: {k: from_torch_named(v) if any(n is not None for n in v.names) else NT(v, ()) for k, v in d.items()}
    prob = Problem(P, Q, nt(inp['data']), inputs=nt(inp['inputs']), params=nt(inp['params']), device="cuda:0")
    isamp = prob.sample(30, reparam=False, seed=3).importance_sample(50, seed=1)
    assert isamp['z'].axes == ('N', 'plate_1') and isamp.N == 50
    g = t.Generator().manual_seed(9)
    M2, N2 = 360, 8
    x2 = (t.rand(M2, N2, 18, generator=g) < 0.107).to(dt)
    x2[:300, :5] = inp['inputs']['x'].rename(None)
    obs2 = (t.rand(M2, N2, generator=g) < 0.5).to(dt)
    obs2[:300, :5] = inp['data']['obs'].rename(None)
    ext_in = {'x': NT(x2, ('plate_1', 'plate_2'))}
    ext = isamp.extend({'plate_1': M2, 'plate_2': N2}, ext_in, seed=4)
    assert ext['z'].t.shape == (50, M2, 18) and ext['obs'].t.shape == (50, M2, N2)
    assert t.equal(ext['z'].t[:, :300], isamp['z'].t)
    assert t.equal(ext['mu_z'].t, isamp['mu_z'].t)
    assert t.equal(ext['obs'].t[:, :300, :5].cpu(), inp['data']['obs'].rename(None).expand(50, 300, 5))
    m = ext.moments([('z', lambda z: z)])[0]
    assert m.t.shape == (M2, 18) and m.axes == ('plate_1',)
    pll = ext.predictive_ll({'obs': NT(obs2, ('plate_1', 'plate_2'))})
    lat = {k: NT(v.t.cpu(), v.axes) for k, v in ext.items() if k != 'obs'}
    ref = PO.predictive_ll(P, lat, {'obs': NT(obs2, ('plate_1', 'plate_2'))}, prob.platesizes, ext_in, 50, dt)
    assert abs(float(pll['obs'].cpu()) - float(ref['obs'])) <= 3e-5 * abs(float(ref['obs']))
    with pytest.raises(Exception, match="extended versions"):
        isamp.extend({'plate_1': M2})
