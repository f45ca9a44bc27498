"""The reference-facing call surface (alan_b200.problem: Problem / Sample mirrors of
src/alan/Problem.py and src/alan/Sample.py) on the GPU against the reference goldens: the same
checks as the reference's own tests/test_problem_vs_itself.py make between its computation
strategies, here between the reference and the B200 engine on identical samples."""
import pytest
import torch as t

import models
from alan_b200 import model as M
from alan_b200.named import NT
from golden_io import load, rel_err, tol, TAGS
from uniforms import UniformSource

pytestmark = pytest.mark.gpu
CASES = list(models.CASES)


def _problem(g, case, requires_grad=False):
    from alan_b200.problem import Problem
    P, Q = models.build(case, M, t.float64 if g['dtype'] == 'torch.float64' else t.float32)
    nt = lambda d, rg=False: {k: NT(v[0].clone().requires_grad_(rg), v[1]) for k, v in d.items()}
    params = nt(g["params"], requires_grad)
    prob = Problem(P, Q, nt(g["data"]), inputs=nt(g["inputs"]), params=params, device="cuda:0")
    return prob, params


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_elbo_vi_backward(case, tag):
    g = load(case, tag)
    prob, params = _problem(g, case, requires_grad=True)
    sample = {k: NT(v[0].clone().requires_grad_(True), v[1]) for k, v in g["sample"].items()}
    s = prob.sample_from(sample, reparam=True)
    L = s.elbo_vi()
    assert L.ndim == 0 and L.is_cuda
    assert rel_err(L.detach().cpu(), g["elbo"]) < tol(tag)
    L.backward()
    for n, ref in g["grad_params"].items():
        assert rel_err(params[n].t.grad, ref) < 30 * tol(tag), n
    for n, ref in g["grad_sample"].items():
        assert rel_err(sample[n].t.grad, ref) < 30 * tol(tag), n
    # RWS: same value, no gradient reaches the samples
    for v in sample.values():
        v.t.grad = None
    L2 = s.elbo_rws()
    assert rel_err(L2.detach().cpu(), g["elbo"]) < tol(tag)
    if L2.requires_grad:
        L2.backward()
    assert all(v.t.grad is None for v in sample.values())
    assert not s.elbo_nograd().requires_grad


def test_elbo_vi_needs_reparam():
    g = load("cfg1_lgl", "f32")
    prob, _ = _problem(g, "cfg1_lgl")
    s = prob.sample_from({k: NT(*v) for k, v in g["sample"].items()}, reparam=False)
    with pytest.raises(Exception, match="reparam"):
        s.elbo_vi()


def test_missing_sample_and_bad_data_raise():
    from alan_b200.problem import Problem
    g = load("cfg1_lgl", "f32")
    prob, _ = _problem(g, "cfg1_lgl")
    with pytest.raises(Exception, match="no sample was provided"):
        prob.sample_from({'a': NT(*g["sample"]['a'])})
    P, Q = models.CASES["cfg1_lgl"][0](M)
    with pytest.raises(Exception, match="no data was provided"):
        Problem(P, Q, {}, device="cuda:0")


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_marginals_moments_ess(case, tag):
    g = load(case, tag)
    prob, _ = _problem(g, case)
    s = prob.sample_from({k: NT(*v) for k, v in g["sample"].items()})
    joints = models.CASES[case][5]
    marg = s.marginals(joints=joints)
    groups = prob.Q.groupvarnames()
    for key, (ref, axes) in g["marginals"].items():
        k2 = tuple(sorted(key, key=groups.index))
        mine = marg.weights[k2].order(axes).t.cpu()
        assert rel_err(mine, ref) < 30 * tol(tag), key
    for grp, e in marg.ess().items():
        w = marg.weights[(grp,)]
        assert (e.t > 0.999).all() and (e.t <= w.named_sizes['K_' + grp] * 1.001).all()
    moms = s.moments([(v, models.MOMENT_FUNCS[f]) for v, f in g["moment_specs"]])
    for mine, (ref, axes) in zip(moms, g["moments"]):
        assert rel_err(mine.order(axes).t.cpu(), ref) < 30 * tol(tag)


@pytest.mark.parametrize("case", [c for c in CASES if models.CASES[c][6] is not None])
def test_importance_sample(case):
    g = load(case, "f32")
    prob, _ = _problem(g, case)
    sample = {k: NT(*v) for k, v in g["sample"].items()}
    s = prob.sample_from(sample)
    N = g["N"]
    run = s._runner(N=N)
    src = UniformSource(g["uniform_seed"], N, g["platesizes"], list(g["platesizes"]))
    us = [src.draw(batch)[0].cuda() for batch, ks in run.comp.plan.sample_steps]
    post = s.importance_sample(N, uniforms=us)
    total = bad = 0
    for grp, (ref, axes) in g["indices"].items():
        mine = s.indices[grp].order(axes).t.cpu()
        total += ref.numel()
        bad += (mine != ref).sum().item()
    assert bad <= 1e-3 * total
    v2g = prob.Q.varname2groupvarname()
    for name, x in sample.items():
        out = post[name]
        assert out.axes[0] == 'N' and out.t.shape[0] == N
        # every posterior sample is one of the K prior samples of the same plate cell
        grp = v2g[name]
        plates = out.axes[1:]
        xk = x.order(plates + ('K_' + grp,)).t.cuda()
        idx = s.indices[grp].order(('N',) + plates).t
        gathered = t.gather(xk.unsqueeze(0).expand(N, *xk.shape), len(plates) + 1,
                            idx.reshape(*idx.shape, 1, *([1] * len(x.pos_shape))).expand(*idx.shape, 1, *x.pos_shape)).squeeze(len(plates) + 1)
        assert t.equal(out.t, gathered)
    # default uniforms: seeded, reproducible
    a = s.importance_sample(N, seed=3)
    b = s.importance_sample(N, seed=3)
    assert all(t.equal(a[k].t, b[k].t) for k in a)


def test_plans_are_cached_per_problem():
    """A new Sample of the same problem and shapes (every training iteration draws one) reuses the compiled plan
    and its device workspace."""
    g = load("cfg2_movielens", "f32")
    prob, _ = _problem(g, "cfg2_movielens", requires_grad=True)
    s1 = prob.sample_from({k: NT(*v) for k, v in g["sample"].items()})
    L1 = s1.elbo_rws()
    n = len(prob._runners)
    s2 = prob.sample_from({k: NT(v[0] * 1.0, v[1]) for k, v in g["sample"].items()})
    L2 = s2.elbo_rws()
    assert len(prob._runners) == n
    assert t.equal(L1.detach(), L2.detach())
