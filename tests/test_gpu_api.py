"""The reference-facing call surface (alan_b200.problem: Problem / Sample mirrors of
src/alan/Problem.py and src/alan/Sample.py) on the GPU against the reference goldens: the same
checks as the reference's own tests/test_problem_vs_itself.py make between its computation
strategies, here between the reference and the B200 engine on identical samples."""
import pytest
import torch as t

import models
from alan_b200 import model as M
from alan_b200.named import NT, from_torch_named
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
    # ESS = 1 / sum_K w^2 (Marginals.py:52-61) against the same formula on the reference's golden marginals
    ess = marg.ess()
    worst = None
    for key, (ref, axes) in g["marginals"].items():
        if len(key) != 1:
            continue
        grp = key[0]
        kd = tuple(i for i, a in enumerate(axes) if a.startswith("K_"))
        ref_ess = 1.0 / (ref.double() ** 2).sum(kd)
        plates = tuple(a for a in axes if not a.startswith("K_"))
        mine = ess[grp].order(plates).t.cpu()
        assert rel_err(mine, ref_ess) < 100 * tol(tag), grp
        worst = ref_ess.min().item() if worst is None else min(worst, ref_ess.min().item())
    assert abs(marg.min_ess() - worst) <= 100 * tol(tag) * worst
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


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_moments_sample_marginal(case, tag):
    """The reference's own `test_moments_sample_marginal` (tests/test_problem_vs_itself.py:71-88): `sample.moments`
    (source-term gradient) and `marginals.moments` (sum_K f(x) w, Marginals.py:31-46) agree on the same sample --
    same rtol 1e-4 / atol 1e-5 -- and both match the reference's golden moments."""
    g = load(case, tag)
    prob, _ = _problem(g, case)
    s = prob.sample_from({k: NT(*v) for k, v in g["sample"].items()})
    specs = [(v, models.MOMENT_FUNCS[f]) for v, f in g["moment_specs"]]
    a = s.moments(specs)
    b = s.marginals().moments(specs)
    for ma, mb, (ref, axes) in zip(a, b, g["moments"]):
        assert ma.axes == mb.axes or set(ma.axes) == set(mb.axes)
        assert t.allclose(ma.t, mb.order(ma.axes).t, rtol=1e-4, atol=1e-5)
        assert rel_err(mb.order(axes).t.cpu(), ref) < 30 * tol(tag)


def test_inline_moment_lambdas_do_not_alias():
    """Two different inline lambdas in a row must not hit each other's cached plan (their id() can be recycled once
    the first is freed; the cache entry keeps the function alive)."""
    g = load("cfg1_lgl", "f32")
    prob, _ = _problem(g, "cfg1_lgl")
    s = prob.sample_from({k: NT(*v) for k, v in g["sample"].items()})
    m1 = s.moments([('a', lambda x: x)])[0].t.clone()
    m2 = s.moments([('a', lambda x: x * x)])[0].t.clone()
    m1b = s.moments([('a', lambda x: x)])[0].t.clone()
    ref = {f: r for (v, f), (r, _) in zip(g["moment_specs"], g["moments"]) if v == 'a'}
    assert rel_err(m1.cpu(), ref['mean']) < 3e-4 and rel_err(m2.cpu(), ref['mean2']) < 3e-4
    assert t.equal(m1, m1b) and not t.equal(m1, m2)


def test_interleaved_forwards_keep_their_own_gradients():
    """`(l1 + l2).backward()` with two samples of one problem: both forwards share a cached runner (one device
    workspace); each backward must see ITS forward's intermediates (the runner recomputes a stale forward)."""
    g = load("cfg2_movielens", "f32")
    grads = []
    for which in (0, 1, 2):
        prob, params = _problem(g, "cfg2_movielens", requires_grad=True)
        s1 = prob.sample_from({k: NT(*v) for k, v in g["sample"].items()})
        s2 = prob.sample_from({k: NT(v[0] * 0.9 + 0.05, v[1]) for k, v in g["sample"].items()})
        if which == 0:
            s1.elbo_rws().backward()
        elif which == 1:
            s2.elbo_rws().backward()
        else:
            l1, l2 = s1.elbo_rws(), s2.elbo_rws()
            assert len(prob._runners) == 1
            (l1 + l2).backward()
        grads.append({k: v.t.grad.clone() for k, v in params.items()})
    for k in grads[0]:
        assert rel_err(grads[2][k], grads[0][k] + grads[1][k]) < 1e-5, k


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case,plate,size", [("cfg1_lgl", "T", 3), ("cfg1_lglp", "T", 4), ("cfg2_movielens", "plate_1", 5),
                                              ("cfg3_radon", "States", 2), ("ref_dangling", "T", 3)])
def test_compstrat_split_matches_single_pass(case, plate, size, tag):
    """The reference's `test_compstrat_elbo_vi / _rws / _moments` (tests/test_problem_vs_itself.py:231-280):
    no_checkpoint == checkpoint == Split(...) on the same sample -- log-evidence, parameter and sample gradients,
    marginals and moments.  Split sizes leave a ragged last block on purpose."""
    from alan_b200.strategy import Split, no_checkpoint, checkpoint
    g = load(case, tag)
    res = {}
    for name, strat in (("none", no_checkpoint), ("ckpt", checkpoint), ("split", Split(plate, size))):
        prob, params = _problem(g, case, requires_grad=True)
        sample = {k: NT(v[0].clone().requires_grad_(True), v[1]) for k, v in g["sample"].items()}
        s = prob.sample_from(sample, reparam=True)
        L = s.elbo_vi(computation_strategy=strat)
        L.backward()
        moms = s.moments([(v, models.MOMENT_FUNCS[f]) for v, f in g["moment_specs"]], computation_strategy=strat)
        marg = s.marginals(computation_strategy=strat)
        res[name] = (L.detach().clone(), {k: v.t.grad.clone() for k, v in params.items()},
                     {k: v.t.grad.clone() for k, v in sample.items()}, [m.t.clone() for m in moms],
                     {k: w.t.clone() for k, w in marg.weights.items()},
                     s.elbo_nograd(computation_strategy=strat).clone())
    tl = tol(tag)
    assert rel_err(res["none"][0].cpu(), g["elbo"]) < tl
    assert t.equal(res["none"][0], res["ckpt"][0])
    a, b = res["none"], res["split"]
    assert rel_err(b[0], a[0]) < tl and rel_err(b[5], a[0]) < tl
    for k in a[1]:
        assert rel_err(b[1][k], a[1][k]) < 30 * tl, k
    for k in a[2]:
        assert rel_err(b[2][k], a[2][k]) < 30 * tl, k
    for x, y in zip(a[3], b[3]):
        assert t.allclose(x, y, rtol=1e-4, atol=1e-5)                   # upstream's own tolerance for this pair
    for k in a[4]:
        assert rel_err(b[4][k], a[4][k]) < 30 * tl, k
    if case == "cfg3_radon":                       # only top-level plates can be split here; a nested one is refused
        with pytest.raises(Exception, match="Split"):
            prob.sample_from({k: NT(*v) for k, v in g["sample"].items()}).elbo_nograd(computation_strategy=Split("Counties", 2))


def test_pipelined_runner_matches_single_steps():
    """engine.PipelinedRunner (two host batches in flight: H2D of s+1 over the kernels of s over the D2H of s-1) returns,
    for every submitted batch, exactly what Runner.step returns for that batch alone."""
    from alan_b200.engine import Compiled, Runner, PipelinedRunner
    P, Q = models.build('cfg2_movielens', M, t.float32)
    inp = models.movielens_inputs(M=64, N=5, dtype=t.float32)
    nt = lambda d: {k: from_torch_named(v) if any(n is not None for n in v.names) else NT(v, ()) for k, v in d.items()}
    ip, data = {**nt(inp['inputs']), **nt(inp['params'])}, nt(inp['data'])
    K, d = 8, 18
    g = t.Generator().manual_seed(4)
    def sample(seed_scale):
        return {'mu_z': NT(seed_scale * t.randn(K, d, generator=g), ('K_mu_z',)),
                'psi_z': NT(0.3 * t.randn(K, d, generator=g), ('K_psi_z',)),
                'z': NT(seed_scale * t.randn(64, K, d, generator=g), ('plate_1', 'K_z'))}
    batches = [sample(0.5 + 0.1 * i) for i in range(7)]
    comp = Compiled(P, Q, batches[0], ip, data, grad_names=list(inp['params']))
    ref = Runner(comp, 'cuda:0')
    want = []
    for b in batches:
        tens = [x.cuda() for x in comp.canonical_inputs(b, ip, data)]
        lp = ref.forward_raw(tens)
        gr = ref.backward_raw(tens)
        want.append((lp.cpu().clone(), {n: v.cpu().clone() for n, v in gr.items()}))
    pipe = PipelinedRunner(comp, 'cuda:0')
    hosts = [pipe.pin(b, ip, data) for b in batches]
    got = []
    tk_prev = None
    for h in hosts:
        tk = pipe.submit(h)
        if tk_prev is not None:
            lp, gr = pipe.result(tk_prev)
            got.append((lp.clone(), {n: v.clone() for n, v in gr.items()}))
        tk_prev = tk
    lp, gr = pipe.result(tk_prev)
    got.append((lp.clone(), {n: v.clone() for n, v in gr.items()}))
    with pytest.raises(Exception, match="not in flight"):
        pipe.result(0)
    for (lw, gw), (lg, gg) in zip(want, got):
        assert t.equal(lw, lg)
        for n in gw:
            assert t.equal(gw[n], gg[n]), n


@pytest.mark.parametrize("dt", [t.float32, t.float64])
@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 4099, 1 << 20])
def test_widen_u8_bit_exact(n, dt):
    """alan_b200_widen_u8 (byte-typed inputs widened on the device) against torch's own cast, ragged tails included."""
    from alan_b200 import runtime
    src = t.randint(0, 256, (n,), dtype=t.uint8, generator=t.Generator().manual_seed(n)).cuda()
    assert t.equal(runtime.widen(src, dtype=dt).cpu(), src.cpu().to(dt))
    b = (src > 127)
    assert t.equal(runtime.widen(b, dtype=dt).cpu(), b.cpu().to(dt))
    with pytest.raises(Exception, match="uint8"):
        runtime.widen(src.float())


def test_byte_typed_inputs_match_float_inputs(monkeypatch):
    """Binary covariates / 0-1 observations handed over as uint8 (PipelinedRunner.pin keeps them one byte per element,
    the device widens them) give bit for bit what the float copies give, through the pipelined entry point and through
    the Problem API."""
    from alan_b200.engine import Compiled, PipelinedRunner
    from alan_b200.problem import Problem
    P, Q = models.build('cfg2_movielens', M, t.float32)
    inp = models.movielens_inputs(M=64, N=5, dtype=t.float32)
    nt = lambda d: {k: from_torch_named(v) if any(n is not None for n in v.names) else NT(v, ()) for k, v in d.items()}
    ip, data = {**nt(inp['inputs']), **nt(inp['params'])}, nt(inp['data'])
    assert set(ip['x'].t.unique().tolist()) <= {0.0, 1.0} and set(data['obs'].t.unique().tolist()) <= {0.0, 1.0}
    ip8 = dict(ip, x=NT(ip['x'].t.to(t.uint8), ip['x'].axes))
    data8 = dict(data, obs=NT(data['obs'].t.to(t.bool), data['obs'].axes))
    K, d = 8, 18
    g = t.Generator().manual_seed(5)
    smp = {'mu_z': NT(0.6 * t.randn(K, d, generator=g), ('K_mu_z',)), 'psi_z': NT(0.3 * t.randn(K, d, generator=g), ('K_psi_z',)),
           'z': NT(0.6 * t.randn(64, K, d, generator=g), ('plate_1', 'K_z'))}
    comp = Compiled(P, Q, smp, ip, data, grad_names=list(inp['params']))
    pipe = PipelinedRunner(comp, 'cuda:0')
    from alan_b200 import engine
    monkeypatch.setattr(engine, "NARROW_MIN_NUMEL", 1)          # small tensors are normally widened on the host
    hf, h8 = pipe.pin(smp, ip, data), pipe.pin(smp, ip8, data8)
    assert sum(x.numel() * x.element_size() for x in h8) < sum(x.numel() * x.element_size() for x in hf)
    assert {x.dtype for x in h8} == {t.float32, t.uint8, t.bool}
    res = []
    for h in (hf, h8, h8, hf, h8):
        lp, gr = pipe.result(pipe.submit(h))
        res.append((lp.clone(), {n: v.clone() for n, v in gr.items()}))
    for lp, gr in res[1:]:
        assert t.equal(lp, res[0][0])
        for n in gr:
            assert t.equal(gr[n], res[0][1][n]), n
    # the Problem API: uint8 inputs / data are moved as bytes and widened by the library's kernel
    names = list(inp['params'])
    out = []
    for i_, d_ in ((ip, data), (ip8, data8)):
        par = {k: NT(i_[k].t.clone().requires_grad_(True), i_[k].axes) for k in names}
        prob = Problem(P, Q, d_, inputs={k: v for k, v in i_.items() if k not in names}, params=par, device='cuda:0')
        L = prob.sample_from(smp).elbo_rws()
        L.backward()
        out.append((L.detach().cpu(), {k: par[k].t.grad.clone() for k in names}))
    assert t.equal(out[0][0], out[1][0])
    for k in names:
        assert t.equal(out[0][1][k], out[1][1][k]), k
