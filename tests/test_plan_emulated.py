"""Planner logic checked on CPU: plan ops are executed by tests/plan_emulator.py (torch
semantics of the CUDA kernels) and compared with the reference goldens.  The same plans run on
the GPU through the C ABI in tests/test_gpu_parity.py."""
import pytest
import torch as t

import models
from alan_b200 import model as M
from alan_b200.engine import Compiled
from alan_b200.named import NT
from golden_io import load, rel_err, tol, TAGS
from plan_emulator import Emu
from uniforms import UniformSource

CASES = list(models.CASES)


def run_fwd_bwd(comp, inputs, want_grads=True):
    plan = comp.plan
    lp = t.zeros(1, dtype=plan.dtype)
    emu = Emu(plan, inputs, outputs={0: lp})
    for seg in plan.programs[:plan.n_fwd]:
        emu.run(seg)
    grads = {}
    if want_grads and plan.n_bwd:
        emu.outputs = {i: t.zeros(plan.input_pts[n].numel, dtype=plan.dtype) for i, n in enumerate(plan.grad_inputs)}
        emu.aux = {0: t.ones(1, dtype=plan.dtype)}
        for seg in plan.programs[plan.n_fwd:plan.n_fwd + plan.n_bwd]:
            emu.run(seg)
        grads = {n: emu.outputs[i].reshape(plan.input_pts[n].shape) for i, n in enumerate(plan.grad_inputs)}
    return lp[0], grads, emu


def grad_as(comp, grads, name, axes):
    pt = comp.plan.input_pts[name]
    return NT(grads[name], pt.axes).order(axes).t if pt.axes else grads[name]


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_elbo_and_grads(case, tag):
    g = load(case, tag)
    P, Q = models.build(case, M, TAGS[tag])
    names = list(g["grad_sample"]) + list(g["grad_params"])
    comp = Compiled(P, Q, g["sample_nt"], g["inputs_params_nt"], g["data_nt"], grad_names=names)
    inputs = comp.canonical_inputs(g["sample_nt"], g["inputs_params_nt"], g["data_nt"])
    lp, grads, _ = run_fwd_bwd(comp, inputs)
    assert rel_err(lp, g["elbo"]) < tol(tag)
    for n in g["grad_sample"]:
        assert rel_err(grad_as(comp, grads, n, g["sample"][n][1]), g["grad_sample"][n]) < 30 * tol(tag), n
    for n in g["grad_params"]:
        assert rel_err(grad_as(comp, grads, n, g["params"][n][1]), g["grad_params"][n]) < 30 * tol(tag), n


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_marginals_and_moments(case, tag):
    g = load(case, tag)
    P, Q = models.build(case, M, TAGS[tag])
    dtype = TAGS[tag]
    g2p = Q.groupvarname2platenames()
    groups = Q.groupvarnames()
    elf = {}
    for key in g["marginals"]:
        gs = tuple(sorted(key, key=groups.index))
        axes = tuple(M.Kname(x) for x in gs) + tuple(g2p[gs[0]])
        sizes = {**{a: s for v in g["sample_nt"].values() for a, s in v.named_sizes.items()}, **g["platesizes"]}
        elf[key] = NT(t.zeros([sizes[a] for a in axes], dtype=dtype), axes)
    moms = [((v,), models.MOMENT_FUNCS[f]) for v, f in g["moment_specs"]]
    comp = Compiled(P, Q, g["sample_nt"], g["inputs_params_nt"], g["data_nt"], extra_log_factors=elf,
                    moment_specs=moms, grad_names=list(elf.keys()))
    inputs = comp.canonical_inputs(g["sample_nt"], g["inputs_params_nt"], g["data_nt"], elf)
    lp, grads, _ = run_fwd_bwd(comp, inputs)
    assert rel_err(lp, g["elbo"]) < tol(tag)
    for key, (ref, axes) in g["marginals"].items():
        name = comp.elf_keys[key]
        pt = comp.plan.input_pts[name]
        mine = NT(grads[name], pt.axes).order(axes).t
        assert rel_err(mine, ref) < 30 * tol(tag), key
    for (jname, plates, pos), (ref, axes) in zip(comp.moment_inputs, g["moments"]):
        mine = NT(grads[jname], plates).order(axes).t
        assert rel_err(mine, ref) < 30 * tol(tag), jname


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", [c for c in CASES if models.CASES[c][6] is not None])
def test_resampling(case, tag):
    g = load(case, tag)
    P, Q = models.build(case, M, TAGS[tag])
    N = g["N"]
    comp = Compiled(P, Q, g["sample_nt"], g["inputs_params_nt"], g["data_nt"], N=N)
    inputs = comp.canonical_inputs(g["sample_nt"], g["inputs_params_nt"], g["data_nt"])
    lp, _, emu = run_fwd_bwd(comp, inputs, want_grads=False)
    plan = comp.plan
    src = UniformSource(g["uniform_seed"], N, g["platesizes"], list(g["platesizes"]))
    aux = {}
    for i, (batch_axes, ks) in enumerate(plan.sample_steps):
        u, axes = src.draw(batch_axes)
        assert axes == tuple(batch_axes) + ('N',)
        aux[i] = u.reshape(-1)
    emu.aux = aux
    emu.outputs = {}
    for gi, (grp, plates) in enumerate(plan.sample_groups):
        n = N
        for a in plates:
            n *= g["platesizes"][a]
        emu.outputs[gi] = t.zeros(n, dtype=t.long)
    emu.run(plan.programs[plan.sample_prog])
    total = bad = 0
    for gi, (grp, plates) in enumerate(plan.sample_groups):
        ref, axes = g["indices"][grp]
        mine = NT(emu.outputs[gi].reshape([N] + [g["platesizes"][a] for a in plates]), ('N',) + tuple(plates))
        mine = mine.order(axes).t
        total += ref.numel()
        bad += (mine != ref).sum().item()
    assert bad <= 1e-3 * total, f"{bad}/{total}"


@pytest.mark.parametrize("dtype", [t.float32, t.float64])
def test_fused_paths_selected_and_match_oracle(dtype):
    """At sizes where the planner picks the tuned kernels (normal_fan fused with its LSE, dot), the
    plan still reproduces the oracle's log-evidence and RWS gradients (emulated semantics)."""
    from oracle import logpq_oracle as O
    from alan_b200.named import from_torch_named
    P, Q = models.movielens_model(M)
    inp = models.movielens_inputs(M=20, N=3, d=18, seed=3, dtype=dtype)
    g = t.Generator().manual_seed(9)
    K = 8
    r = lambda *s: (0.7 * t.randn(s, generator=g, dtype=t.float64)).to(dtype)
    sample = {'mu_z': NT(r(K, 18), ('K_mu_z',)), 'psi_z': NT(r(K, 18) - 0.5, ('K_psi_z',)),
              'z': NT(r(20, K, 18), ('plate_1', 'K_z'))}
    ip = {k: from_torch_named(v) for k, v in {**inp['inputs'], **inp['params']}.items()}
    data = {k: from_torch_named(v) for k, v in inp['data'].items()}
    names = list(inp['params'])
    comp = Compiled(P, Q, sample, ip, data, grad_names=names)
    kinds = [type(op).__name__ for prog in comp.plan.programs for op in prog]
    assert 'FanLseOp' in kinds and 'FanLseBwdOp' in kinds and 'BernDotSumOp' in kinds
    comp_slow = Compiled(P, Q, sample, ip, data, grad_names=names, fast_paths=False)
    assert 'FanLseOp' not in [type(op).__name__ for prog in comp_slow.plan.programs for op in prog]
    inputs = comp.canonical_inputs(sample, ip, data)
    lp, grads, _ = run_fwd_bwd(comp, inputs)
    ipg = {k: NT(v.t.clone().requires_grad_() if k in names else v.t, v.axes) for k, v in ip.items()}
    ref = O.elbo(P, Q, sample, ipg, data)
    rg = t.autograd.grad(ref, [ipg[k].t for k in names])
    tl = 1e-5 if dtype == t.float32 else 1e-10
    assert rel_err(lp, ref) < tl
    for k, rr in zip(names, rg):
        assert rel_err(grad_as(comp, grads, k, ipg[k].axes), rr) < 30 * tl, k
    # a VI-style request (gradient w.r.t. the sample) must fall back to the materialised factor
    comp_vi = Compiled(P, Q, sample, ip, data, grad_names=names + ['z', 'mu_z', 'psi_z'])
    kinds = [type(op).__name__ for prog in comp_vi.plan.programs for op in prog]
    assert 'FanLseOp' not in kinds and 'NormalFanOp' in kinds and 'DotOp' in kinds and 'NormalFanBwdOp' in kinds
    vnames = names + ['z', 'mu_z', 'psi_z']
    lp_vi, grads_vi, _ = run_fwd_bwd(comp_vi, comp_vi.canonical_inputs(sample, ip, data))
    sg = {k: NT(v.t.clone().requires_grad_(), v.axes) for k, v in sample.items()}
    ipg = {k: NT(v.t.clone().requires_grad_() if k in names else v.t, v.axes) for k, v in ip.items()}
    ref = O.elbo(P, Q, sg, ipg, data)
    rg = t.autograd.grad(ref, [ipg[k].t for k in names] + [sg[k].t for k in ('z', 'mu_z', 'psi_z')])
    assert rel_err(lp_vi, ref) < tl
    for k, rr in zip(vnames, rg):
        axes = ipg[k].axes if k in ipg else sg[k].axes
        assert rel_err(grad_as(comp_vi, grads_vi, k, axes), rr) < 30 * tl, k


@pytest.mark.parametrize("K,NG,M_", [(12, 1, 16), (20, 2, 16), (12, 1, 260)])
def test_dense_fan_layout_emulated_matches_oracle(K, NG, M_):
    """fp32 plans whose fused contraction fits csrc/fan_tc2.cuh commit the adjoint to the compact gS layout
    [users, NG fan-group partials, kappa] (plan.py dense_fan_geometry) and, for long plates, the plate sum to the
    kernel's partial rows; the emulated plan with those layouts still reproduces the oracle's log-evidence and RWS
    gradients."""
    from oracle import logpq_oracle as O
    from alan_b200.named import from_torch_named
    dtype = t.float32
    P, Q = models.movielens_model(M)
    inp = models.movielens_inputs(M=M_, N=3, d=18, seed=5, dtype=dtype)
    g = t.Generator().manual_seed(11)
    r = lambda *s: (0.7 * t.randn(s, generator=g, dtype=t.float64)).to(dtype)
    sample = {'mu_z': NT(r(K, 18), ('K_mu_z',)), 'psi_z': NT(r(K, 18) - 0.5, ('K_psi_z',)),
              'z': NT(r(M_, K, 18), ('plate_1', 'K_z'))}
    ip = {k: from_torch_named(v) for k, v in {**inp['inputs'], **inp['params']}.items()}
    data = {k: from_torch_named(v) for k, v in inp['data'].items()}
    names = list(inp['params'])
    comp = Compiled(P, Q, sample, ip, data, grad_names=names)
    fan = [op for prog in comp.plan.programs for op in prog if type(op).__name__ == 'FanLseOp'][0]
    assert fan.dense is not None and fan.dense[2] == NG
    bwd = [op for prog in comp.plan.programs for op in prog if type(op).__name__ == 'FanLseBwdOp'][0]
    assert bwd.gS.numel == M_ * NG * K
    # a plate long enough to be summed in two stages rides in the dense kernel's epilogue (fused plate sum)
    assert (fan.psum is not None) == (M_ >= 256)
    lp, grads, _ = run_fwd_bwd(comp, comp.canonical_inputs(sample, ip, data))
    ipg = {k: NT(v.t.clone().requires_grad_() if k in names else v.t, v.axes) for k, v in ip.items()}
    ref = O.elbo(P, Q, sample, ipg, data)
    rg = t.autograd.grad(ref, [ipg[k].t for k in names])
    assert rel_err(lp, ref) < 1e-5
    for k, rr in zip(names, rg):
        assert rel_err(grad_as(comp, grads, k, ipg[k].axes), rr) < 3e-4, k


def test_importance_sample_through_timeseries_raises_like_the_reference():
    """`sample_Ks_timeseries` is unfinished upstream (README.md:41-44; IndexError at
    reduce_Ks.py:223 on torch 2.11), so a resampling program through a Timeseries is refused when the plan is built,
    with the reference's own explanation, instead of producing indices nobody can check."""
    from golden_io import load
    g = load("cfg4_timeseries", "f32")
    P, Q = models.build("cfg4_timeseries", M, t.float32)
    Compiled(P, Q, g["sample_nt"], g["inputs_params_nt"], g["data_nt"])           # log-evidence plan: fine
    with pytest.raises(Exception, match="Timeseries is unfinished in the reference"):
        Compiled(P, Q, g["sample_nt"], g["inputs_params_nt"], g["data_nt"], N=5)


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_marginals_moments_equal_sample_moments(case, tag):
    """The reference's own pin `test_moments_sample_marginal` (tests/test_problem_vs_itself.py:71-88):
    `marginals.moments` (sum_K f(x) w, Marginals.py:31-46 / moments.py:16-35 -- here plan.weighted_moment_plan, run by
    the emulator) equals `sample.moments` (the source-term gradient) on the same sample: rtol 1e-4, atol 1e-5 as
    upstream, against the reference's golden moments and marginals."""
    from alan_b200.plan import weighted_moment_plan, TensorSig
    g = load(case, tag)
    P, Q = models.build(case, M, TAGS[tag])
    groups, v2g = Q.groupvarnames(), Q.varname2groupvarname()
    canon = list(P.all_platenames()) + [M.Kname(x) for x in groups]
    order = lambda axes: tuple(a for a in canon if a in axes)
    for (var, fname), (ref, axes) in zip(g["moment_specs"], g["moments"]):
        x = g["sample_nt"][var]
        w_t, w_axes = g["marginals"][(v2g[var],)]
        w = NT(w_t, w_axes)
        sizes = {**x.named_sizes, **w.named_sizes}
        xa, wa = order(x.axes), order(w.axes)
        plan = weighted_moment_plan({var: TensorSig('sample', xa, x.pos_shape)}, wa, models.MOMENT_FUNCS[fname],
                                    sizes, TAGS[tag], canon)
        ins = []
        for name in plan.input_names:
            if name in plan.const_inputs:
                ins.append(plan.const_inputs[name])
            elif name == '__w':
                ins.append(w.order(wa).t.to(TAGS[tag]).contiguous())
            else:
                ins.append(x.order(xa).t.to(TAGS[tag]).contiguous())
        n_out = 1
        for s_ in plan.out_shape:
            n_out *= s_
        out = t.zeros(max(n_out, 1), dtype=TAGS[tag])
        emu = Emu(plan, ins, outputs={0: out})
        emu.run(plan.programs[0])
        mine = NT(out.reshape(plan.out_shape), plan.out_axes).order(axes).t if plan.out_axes else out.reshape(plan.out_shape)
        assert t.allclose(mine.double(), ref.double(), rtol=1e-4, atol=1e-5), (var, fname)


def test_split_sizes_follow_the_reference_rule():
    """strategy.Split.sizes = SplitDims (Split.py:84-95): [s, ..., s, rem], one element stolen when rem == 1."""
    from alan_b200.strategy import Split, resolve, no_checkpoint, checkpoint
    assert Split('p', 20).sizes(300) == [20] * 15
    assert Split('p', 20).sizes(305) == [20] * 15 + [5]
    assert Split('p', 20).sizes(301) == [20] * 14 + [19, 2]
    assert Split('p', 2).sizes(5) == [2, 2, 1]
    with pytest.raises(AssertionError):
        Split('p', 10).sizes(10)
    assert resolve(None) is None and resolve(no_checkpoint) is None and resolve(checkpoint) is None
    sp = Split('plate_1', 7)
    assert resolve(sp) is sp
    with pytest.raises(Exception, match="computation_strategy"):
        resolve("nope")


@pytest.mark.parametrize("dtype", [t.float32, t.float64])
@pytest.mark.parametrize("family", ["OneHotCategorical", "Multinomial", "Categorical", "RelaxedOneHotCategorical"])
def test_vector_families_vs_oracle(family, dtype):
    """OneHotCategorical / Multinomial (densities composed from the VM's primitive operations, plan.py COMPOSED) through
    the plan emulator against torch.distributions + autograd in the oracle."""
    from oracle import logpq_oracle as O
    P, Q, sample, params, data = models.vector_family_case(M, family, dtype)
    names = ['p', 's'] + list(params)
    comp = Compiled(P, Q, sample, params, data, grad_names=names)
    lp, grads, _ = run_fwd_bwd(comp, comp.canonical_inputs(sample, params, data))
    sg = {k: NT(v.t.clone().requires_grad_(), v.axes) for k, v in sample.items()}
    pg = {k: NT(v.t.clone().requires_grad_(), v.axes) for k, v in params.items()}
    ref = O.elbo(P, Q, sg, pg, data)
    rg = t.autograd.grad(ref, [sg['p'].t, sg['s'].t] + [pg[k].t for k in params], allow_unused=True)
    tl = 2e-5 if dtype == t.float32 else 1e-10
    assert t.isfinite(ref) and rel_err(lp, ref) < tl
    for k, rr in zip(names, rg):
        assert rel_err(grads[k].reshape(rr.shape), rr) < 50 * tl, k
    with pytest.raises(Exception, match="total_count"):
        M.Multinomial(3, probs='p')


COMPOSED_SCALAR = {
    'Gumbel': (lambda ns: ns.Gumbel('a', lambda b: b.exp()), lambda r, n: r(n)),
    'Weibull': (lambda ns: ns.Weibull(lambda a: a.exp(), lambda b: b.exp() + 0.5), lambda r, n: r(n).abs() + 0.2),
    'Pareto': (lambda ns: ns.Pareto(0.1, lambda a: a.exp() + 0.5), lambda r, n: r(n).abs() + 0.2),
    'HalfCauchy': (lambda ns: ns.HalfCauchy(lambda a: a.exp()), lambda r, n: r(n).abs() + 0.1),
    'Chi2': (lambda ns: ns.Chi2(lambda a: a.exp() + 1.0), lambda r, n: r(n).abs() + 0.2),
    'Geometric': (lambda ns: ns.Geometric(logits='a'), lambda r, n: (r(n).abs() * 3).floor()),
    'Kumaraswamy': (lambda ns: ns.Kumaraswamy(lambda a: a.exp() + 0.5, lambda b: b.exp() + 0.5), lambda r, n: r(n).sigmoid()),
    'FisherSnedecor': (lambda ns: ns.FisherSnedecor(lambda a: a.exp() + 2.0, 5.0), lambda r, n: r(n).abs() + 0.2),
    'RelaxedBernoulli': (lambda ns: ns.RelaxedBernoulli(lambda b: b.exp() + 0.3, probs=lambda a: a.sigmoid()),
                         lambda r, n: r(n).sigmoid()),
    'ContinuousBernoulli': (lambda ns: ns.ContinuousBernoulli(probs=lambda a: a.sigmoid()), lambda r, n: r(n).sigmoid()),
    'ContinuousBernoulli_taylor': (lambda ns: ns.ContinuousBernoulli(logits=lambda a: 0.002 * a), lambda r, n: r(n).sigmoid()),
    # concentrations on both sides of 3.75: the small and the large branch of log I0
    'VonMises': (lambda ns: ns.VonMises('a', lambda b: 3.0 + 4.0 * b.exp()), lambda r, n: 2.0 * r(n).tanh()),
    'VonMises_small': (lambda ns: ns.VonMises(0.3, lambda b: b.exp()), lambda r, n: 2.0 * r(n).tanh()),
}


@pytest.mark.parametrize("family", list(COMPOSED_SCALAR))
def test_composed_scalar_families_vs_oracle(family):
    """The scalar families whose density is composed from VM primitives (plan.py COMPOSED), float64, through the plan
    emulator against torch.distributions + autograd (the GPU twin, both dtypes: test_density_families_vs_oracle)."""
    import zlib
    from oracle import logpq_oracle as O
    like, gen = COMPOSED_SCALAR[family]
    P = M.Plate(a=M.Normal(0., 1.), b=M.Normal(-0.3, 0.5), T=M.Plate(y=like(M)))
    Q = M.Plate(a=M.Normal('a_loc', lambda a_ls: a_ls.exp()), b=M.Normal('b_loc', lambda b_ls: b_ls.exp()),
                T=M.Plate(y=M.Data()))
    g = t.Generator().manual_seed(zlib.crc32(family.encode()) % 1000)
    r = lambda *s: t.randn(s, generator=g, dtype=t.float64)
    data = {'y': NT(gen(r, 11), ('T',))}
    params = {'a_loc': NT(0.1 * r(), ()), 'a_ls': NT(-0.5 + 0.1 * r(), ()), 'b_loc': NT(-0.3 + 0.1 * r(), ()),
              'b_ls': NT(-0.7 + 0.1 * r(), ())}
    sample = {'a': NT(0.6 * r(4), ('K_a',)), 'b': NT(-0.3 + 0.4 * r(4), ('K_b',))}
    names = ['a', 'b'] + list(params)
    comp = Compiled(P, Q, sample, params, data, grad_names=names)
    lp, grads, _ = run_fwd_bwd(comp, comp.canonical_inputs(sample, params, data))
    sg = {k: NT(v.t.clone().requires_grad_(), v.axes) for k, v in sample.items()}
    pg = {k: NT(v.t.clone().requires_grad_(), v.axes) for k, v in params.items()}
    ref = O.elbo(P, Q, sg, pg, data)
    rg = t.autograd.grad(ref, [sg['a'].t, sg['b'].t] + [pg[k].t for k in params], allow_unused=True)
    assert t.isfinite(ref) and rel_err(lp, ref) < 1e-10
    for k, rr in zip(names, rg):
        if rr is None:
            assert float(grads[k].abs().max()) == 0.0, k
            continue
        assert rel_err(grads[k].reshape(rr.shape), rr) < 1e-8, k
