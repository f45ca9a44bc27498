"""Global importance sampling (`SampleNonMP`) on the GPU through the public mirror
(`Problem.sample_nonmp` / `SampleNonMP`), against the goldens of the unmodified reference
(tests/golden/make_golden_nonmp.py) and, at BASELINE cfg-2 size, against the oracle (oracle/nonmp_oracle.py)."""
import pytest
import torch as t

import models
from alan_b200 import model as M
from alan_b200.named import NT, from_torch_named
from alan_b200.plan import NONMP_K
from golden_io import TAGS, rel_err, tol
from test_nonmp_cpu import load, CASES, unified_axes

pytestmark = pytest.mark.gpu


def problem_of(g, case, tag, grad=True):
    from alan_b200.problem import Problem
    P, Q = models.build(case, M, TAGS[tag])
    nt = lambda d: {k: NT(v[0].clone(), v[1]) for k, v in d.items()}
    params = nt(g["params"])
    for k, v in params.items():
        v.t.requires_grad_(grad and k in g["grad_params"])
    return Problem(P, Q, nt(g["data"]), inputs=nt(g["inputs"]), params=params, device="cuda:0",
                   platesizes=g["platesizes"])


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_elbo_grads_moments_vs_reference_golden(case, tag):
    from alan_b200.nonmp import SampleNonMP
    g = load(case, tag)
    prob = problem_of(g, case, tag)
    smp = {k: NT(v.t.clone().requires_grad_(k in g["grad_sample"]), v.axes) for k, v in g["sample_nt"].items()}
    s = SampleNonMP(prob, smp, reparam=True)
    assert s.K == g["K"]
    L = s.elbo_vi()
    assert rel_err(L.cpu(), g["elbo"]) < tol(tag)
    L.backward()
    for n in g["grad_sample"]:
        mine = NT(smp[n].t.grad, unified_axes(smp[n].axes)).order(unified_axes(g["sample"][n][1])).t
        assert rel_err(mine, g["grad_sample"][n]) < 30 * tol(tag), n
    for n in g["grad_params"]:
        assert rel_err(prob.params[n].t.grad, g["grad_params"][n]) < 30 * tol(tag), n
    assert rel_err(s.elbo_nograd().cpu(), g["elbo"]) < tol(tag)
    moms = [(v, models.MOMENT_FUNCS[f]) for v, f in g["moment_specs"]]
    for mine, (ref, axes) in zip(s.moments(moms), g["moments"]):
        assert rel_err(mine.order(axes).t.cpu(), ref) < 30 * tol(tag)


@pytest.mark.parametrize("case", ['cfg1_lglp', 'cfg3_radon', 'cfg2_movielens'])
def test_importance_sample_indices_vs_oracle(case):
    """float64: the categorical draw over K from explicit uniforms is bit-equal to the oracle's inverse-CDF rule on the
    reference's lpq, and every latent is gathered at those indices."""
    from alan_b200.nonmp import SampleNonMP
    from oracle import nonmp_oracle as NO
    tag = 'f64'
    g = load(case, tag)
    prob = problem_of(g, case, tag, grad=False)
    s = SampleNonMP(prob, g["sample_nt"], reparam=False)
    N = 200
    u = t.rand(N, dtype=t.float64, generator=t.Generator().manual_seed(3))
    isamp = s.importance_sample(N, uniforms=u.cuda())
    P, Q = models.build(case, M, TAGS[tag])
    ref = NO.importance_sample_idxs(P, Q, g["sample_nt"], g["inputs_params_nt"], g["data_nt"], u)
    assert t.equal(s.indices.t.cpu(), ref)
    for k, v in g["sample_nt"].items():
        kax = [a for a in v.axes if a.startswith('K_')][0]
        plates = tuple(a for a in v.axes if a != kax)
        want = v.order((kax,) + plates).t[ref]
        assert isamp[k].axes == ('N',) + plates
        assert t.equal(isamp[k].t.cpu(), want), k


def test_sample_nonmp_at_cfg2_size_vs_oracle():
    """Problem.sample_nonmp(K) at BASELINE cfg-2 size (300 x 5, d = 18, K = 30): the draw is K independent joint
    samples (IndependentSampler), and elbo / moments equal the oracle's on the drawn sample."""
    from alan_b200.problem import Problem
    from oracle import nonmp_oracle as NO
    dt = t.float32
    inp = models.movielens_inputs(dtype=dt)
    P, Q = models.build('cfg2_movielens', M, dt)
    nt = lambda d: {k: from_torch_named(v) if any(n is not None for n in v.names) else NT(v, ()) for k, v in d.items()}
    prob = Problem(P, Q, nt(inp['data']), inputs=nt(inp['inputs']), params=nt(inp['params']), device="cuda:0")
    s = prob.sample_nonmp(30, reparam=False, seed=5)
    assert s.K == 30 and s.sample['z'].axes[0] == NONMP_K
    cpu = {k: NT(v.t.detach().cpu(), tuple('K_x' if a == NONMP_K else a for a in v.axes)) for k, v in s.sample.items()}
    ip = {**nt(inp['inputs']), **nt(inp['params'])}
    ref = NO.elbo(P, Q, cpu, ip, nt(inp['data']))
    assert rel_err(s.elbo_nograd().cpu(), ref) < 1e-5
    f = lambda z: z
    mine = s.moments([('z', f)])[0]
    want = NO.moments(P, Q, cpu, ip, nt(inp['data']), [(('z',), f)])[0]
    assert rel_err(mine.order(want.axes).t.cpu(), want.t) < 3e-4
    isamp = s.importance_sample(100, seed=2)
    assert isamp['z'].t.shape == (100, 300, 18) and isamp['mu_z'].t.shape == (100, 18)
