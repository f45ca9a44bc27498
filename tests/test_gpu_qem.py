"""QEM on the GPU: `Sample.update_qem_params(lr)` through the public mirror -- posterior moments
from the engine, moving average + mean -> conventional conversion by `alan_b200_qem_update` (csrc/qem.cuh) -- against
the states of the UNMODIFIED reference (tests/golden/make_golden_qem.py), and the conversion kernel alone against the
oracle (oracle/qem_oracle.py) for every family."""
import pytest
import torch as t

import models
from alan_b200 import model as M
from alan_b200.named import NT
from golden_io import TAGS, elem_err
from test_qem_cpu import load, nts, build, CASES

pytestmark = pytest.mark.gpu


def check(prob, ref, tag, what):
    tol = 3e-4 if tag == 'f32' else 1e-9          # fp32: moments of K <= 12 particles, then up to 11 Newton steps
    params, means = prob.qem_params(), prob.qem_means()
    want_p = {**ref['P']['params'], **ref['Q']['params']}
    want_m = {**ref['P']['means'], **ref['Q']['means']}
    assert set(params) == set(want_p) and set(means) == set(want_m), what
    for k, (x, axes) in want_p.items():
        assert elem_err(params[k].order(axes).t.cpu(), x) < tol, (what, 'param', k)
    for k, (x, axes) in want_m.items():
        assert elem_err(means[k].order(axes).t.cpu(), x) < tol, (what, 'mean', k)


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_update_qem_params_vs_reference(case, tag):
    from alan_b200.problem import Problem
    g = load(case, tag)
    P, Q = build(case, TAGS[tag])
    prob = Problem(P, Q, nts(g['data']), params=nts(g['params']), device="cuda:0", platesizes=g['platesizes'])
    check(prob, g['states'][0], tag, 'initial')
    for k, (x, axes) in g['opt_params'].items():
        assert prob.params[k].axes == axes and t.equal(prob.params[k].t.cpu(), x), k
    s = prob.sample_from(nts(g['sample']))
    for step, lr in enumerate(g['lrs'], start=1):
        s.update_qem_params(lr)
        check(prob, g['states'][step], tag, f'after update {step}')
    # the updated parameters are what the next evaluation reads
    assert t.isfinite(s.elbo_nograd()).item()


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("family", ['Normal', 'Bernoulli', 'Poisson', 'Exponential', 'HalfNormal', 'Gamma', 'Beta'])
def test_qem_update_kernel_vs_oracle(family, tag):
    from alan_b200 import runtime
    from oracle import qem_oracle as QO
    dt = TAGS[tag]
    g = t.Generator().manual_seed(7)
    n = 5000
    r = lambda: t.rand(n, generator=g, dtype=t.float64)
    # random conventional parameters -> exact mean parameters (so the conversion has a well-posed answer)
    conv = {'Normal': dict(loc=4 * r() - 2, scale=0.2 + 2 * r()), 'Bernoulli': dict(probs=r()), 'Poisson': dict(rate=0.1 + 5 * r()),
            'Exponential': dict(rate=0.1 + 5 * r()), 'HalfNormal': dict(scale=0.1 + 3 * r()),
            'Gamma': dict(concentration=0.3 + 6 * r(), rate=0.2 + 4 * r()),
            'Beta': dict(concentration1=0.3 + 6 * r(), concentration0=0.3 + 6 * r())}[family]
    new = [m.to(dt) for m in QO.conv2mean(family, conv)]
    # the running means come from perturbed parameters: the mean-parameter space is convex, so the average is valid
    old = [m.to(dt) for m in QO.conv2mean(family, {k: v * (0.8 + 0.4 * r()) if family != 'Bernoulli' else v * r()
                                                    for k, v in conv.items()})]
    lr = 0.3
    want_means = [o.clone().mul_(1 - lr).add_(m, alpha=lr) for o, m in zip(old, new)]
    want = QO.mean2conv(family, want_means)
    means = [o.clone().cuda() for o in old]
    params = [t.empty(n, dtype=dt, device='cuda') for _ in want]
    runtime.qem_update(family, lr, [m.cuda() for m in new], means, params)
    tol = 2e-5 if tag == 'f32' else 1e-10
    for a, b in zip(means, want_means):
        assert elem_err(a.cpu(), b) < (1e-6 if tag == 'f32' else 1e-14)
    from alan_b200.qem import CONV_ARGS
    for arg, got in zip(CONV_ARGS[family], params):
        ok = t.isfinite(want[arg])
        assert ok.double().mean() > 0.999
        assert elem_err(got.cpu()[ok], want[arg][ok]) < tol, arg


def test_sample_nonmp_update_qem_params():
    """SampleNonMP.update_qem_params (SampleNonMP.py:121-125): the same update from global importance weights; against
    the oracle's non-MP moments + conversion."""
    from alan_b200.problem import Problem
    from oracle import nonmp_oracle as NO, qem_oracle as QO
    from alan_b200.qem import bind
    tag, case = 'f64', 'qem_model1'
    g = load(case, tag)
    P, Q = build(case, TAGS[tag])
    prob = Problem(P, Q, nts(g['data']), params=nts(g['params']), device="cuda:0", platesizes=g['platesizes'])
    s = prob.sample_nonmp(16, reparam=False, seed=11)
    cpu = {k: NT(v.t.detach().cpu(), tuple('K_x' if a == 'K_' else a for a in v.axes)) for k, v in s.sample.items()}
    ip0 = {k: NT(v.t.detach().cpu().clone(), v.axes) for k, v in prob.inputs_params().items()}
    mom = NO.moments(prob.P, prob.Q, cpu, ip0, nts(g['data']), [(('a',), QO.MOMENT_FUNCS['mean']), (('a',), QO.MOMENT_FUNCS['mean2'])])
    means0 = {k: v.t.cpu().clone() for k, v in prob.qem_means().items()}
    s.update_qem_params(0.25)
    m1 = means0['a_mean'] * 0.75 + 0.25 * mom[0].t
    m2 = means0['a_mean2'] * 0.75 + 0.25 * mom[1].t
    want = QO.mean2conv('Normal', [m1, m2])
    assert elem_err(prob.qem_params()['a_loc'].t.cpu(), want['loc']) < 1e-9
    assert elem_err(prob.qem_params()['a_scale'].t.cpu(), want['scale']) < 1e-9
