"""Model definitions shared by the golden generator, the oracle tests and the GPU parity tests.

Every builder takes a namespace `ns` offering Plate / Group / Data / Timeseries / Normal /
Bernoulli ... so that the SAME source text builds the model either with the reference
(`ns = alan`, only in the build container) or with the declarative mirror
(`ns = alan_b200.model`, everywhere).  Q parameters are referred to by name (they arrive in
`inputs_params`), which is the reference's `extra_opt_params` mechanism
(/root/reference/src/alan/BoundPlate.py:49-53), so parameter gradients are part of the goldens.

Shapes follow BASELINE.json `configs`:
  cfg1  /root/reference/tests/linear_gaussian_latents.py            (mixture-Q path, Split)
  cfg2  /root/reference/examples/models/movielens/movielens.py:39-74 (MovieLens-shaped)
  cfg3  /root/reference/examples/models/radon/radon.py:62-102 + HMC/radon (nested plates, Group, K^4 joint)
  cfg4  /root/reference/tests/timeseries.py:15-30                    (Timeseries chain)
"""
import math

import torch as t


# --------------------------------------------------------------------------- cfg1
def lgl_model(ns):
    """tests/linear_gaussian_latents.py:28-45 verbatim structure."""
    P = ns.Plate(
        a=ns.Normal(2, 2),
        T=ns.Plate(
            z=ns.Normal('a', 1.3),
            d=ns.Normal('z', 1.5),
        ),
    )
    Q = ns.Plate(
        a=ns.Normal(1, 4),
        T=ns.Plate(
            z=ns.Normal(lambda a: 1.5 * a, 3.5),
            d=ns.Data(),
        ),
    )
    return P, Q


def lgl_inputs(T=10, seed=0, dtype=t.float32):
    g = t.Generator().manual_seed(seed)
    data = {'d': (1.5 + t.randn(T, generator=g, dtype=t.float64)).to(dtype).refine_names('T')}
    return dict(platesizes={'T': T}, data=data, inputs={}, params={})


def lglp_model(ns):
    """cfg1 with named Q parameters so that parameter gradients are exercised."""
    P = ns.Plate(
        a=ns.Normal(2, 2),
        T=ns.Plate(
            z=ns.Normal('a', 1.3),
            d=ns.Normal('z', 1.5),
        ),
    )
    Q = ns.Plate(
        a=ns.Normal('qa_loc', lambda qa_ls: qa_ls.exp()),
        T=ns.Plate(
            z=ns.Normal(lambda a, qz_w, qz_b: qz_w * a + qz_b, lambda qz_ls: qz_ls.exp()),
            d=ns.Data(),
        ),
    )
    return P, Q


def lglp_inputs(T=10, seed=0, dtype=t.float32):
    g = t.Generator().manual_seed(seed)
    r = lambda *s: t.randn(s, generator=g, dtype=t.float64).to(dtype)
    data = {'d': (1.5 + r(T)).refine_names('T')}
    params = {
        'qa_loc': t.tensor(1., dtype=dtype), 'qa_ls': t.tensor(math.log(4.), dtype=dtype),
        'qz_w': (1.5 + 0.1 * r(T)).refine_names('T'), 'qz_b': (0.1 * r(T)).refine_names('T'),
        'qz_ls': t.tensor(math.log(3.5), dtype=dtype),
    }
    return dict(platesizes={'T': T}, data=data, inputs={}, params=params)


# --------------------------------------------------------------------------- cfg2
def movielens_model(ns, d=18):
    P = ns.Plate(
        mu_z=ns.Normal(t.zeros((d,)), t.ones((d,))),
        psi_z=ns.Normal(t.zeros((d,)), t.ones((d,))),
        plate_1=ns.Plate(
            z=ns.Normal("mu_z", lambda psi_z: psi_z.exp()),
            plate_2=ns.Plate(
                obs=ns.Bernoulli(logits=lambda z, x: z @ x),
            ),
        ),
    )
    Q = ns.Plate(
        mu_z=ns.Normal("mu_z_loc", lambda mu_z_ls: mu_z_ls.exp()),
        psi_z=ns.Normal("psi_z_loc", lambda psi_z_ls: psi_z_ls.exp()),
        plate_1=ns.Plate(
            z=ns.Normal("z_loc", lambda z_ls: z_ls.exp()),
            plate_2=ns.Plate(
                obs=ns.Data(),
            ),
        ),
    )
    return P, Q


def movielens_inputs(M=300, N=5, d=18, seed=0, dtype=t.float32):
    g = t.Generator().manual_seed(seed)
    r = lambda *s: t.randn(s, generator=g, dtype=t.float64).to(dtype)
    x = (t.rand(M, N, d, generator=g) < 0.107).to(dtype)          # real features have mean 0.107
    z_true = 0.5 * r(M, d)
    obs = (t.rand(M, N, generator=g) < t.sigmoid((z_true[:, None, :].double() * x.double()).sum(-1))).to(dtype)
    params = {
        'mu_z_loc': 0.1 * r(d), 'mu_z_ls': -0.5 + 0.1 * r(d),
        'psi_z_loc': 0.1 * r(d), 'psi_z_ls': -0.5 + 0.1 * r(d),
        'z_loc': (0.1 * r(M, d)).refine_names('plate_1', None),
        'z_ls': (-0.5 + 0.1 * r(M, d)).refine_names('plate_1', None),
    }
    return dict(platesizes={'plate_1': M, 'plate_2': N},
                data={'obs': obs.refine_names('plate_1', 'plate_2')},
                inputs={'x': x.refine_names('plate_1', 'plate_2', None)},
                params=params)


# --------------------------------------------------------------------------- cfg3
def radon_model(ns):
    """States x Counties x Zips hierarchical radon (HMC/radon/radon.py:24-40 in alan form);
    the global pair is a Group in Q (radon/radon.py:85-88)."""
    P = ns.Plate(
        global_mean=ns.Normal(0., 1.),
        global_log_sigma=ns.Normal(0., 1.),
        States=ns.Plate(
            State_mean=ns.Normal('global_mean', lambda global_log_sigma: global_log_sigma.exp()),
            State_log_sigma=ns.Normal(0., 1.),
            Counties=ns.Plate(
                County_mean=ns.Normal('State_mean', lambda State_log_sigma: State_log_sigma.exp()),
                County_log_sigma=ns.Normal(0., 1.),
                Beta_u=ns.Normal(0., 1.),
                Beta_basement=ns.Normal(0., 1.),
                Zips=ns.Plate(
                    obs=ns.Normal(
                        lambda County_mean, basement, log_uranium, Beta_basement, Beta_u:
                            County_mean + basement * Beta_basement + log_uranium * Beta_u,
                        lambda County_log_sigma: County_log_sigma.exp()),
                ),
            ),
        ),
    )
    Q = ns.Plate(
        global_latents=ns.Group(
            global_mean=ns.Normal('gm_loc', lambda gm_ls: gm_ls.exp()),
            global_log_sigma=ns.Normal('gls_loc', lambda gls_ls: gls_ls.exp()),
        ),
        States=ns.Plate(
            State_mean=ns.Normal('sm_loc', lambda sm_ls: sm_ls.exp()),
            State_log_sigma=ns.Normal('sls_loc', lambda sls_ls: sls_ls.exp()),
            Counties=ns.Plate(
                County_mean=ns.Normal('cm_loc', lambda cm_ls: cm_ls.exp()),
                County_log_sigma=ns.Normal('cls_loc', lambda cls_ls: cls_ls.exp()),
                Beta_u=ns.Normal('bu_loc', lambda bu_ls: bu_ls.exp()),
                Beta_basement=ns.Normal('bb_loc', lambda bb_ls: bb_ls.exp()),
                Zips=ns.Plate(
                    obs=ns.Data(),
                ),
            ),
        ),
    )
    return P, Q


def radon_inputs(S=7, C=10, Z=10, seed=0, dtype=t.float32):
    g = t.Generator().manual_seed(seed)
    r = lambda *s: t.randn(s, generator=g, dtype=t.float64).to(dtype)
    basement = (t.rand(S, C, Z, generator=g) < 0.65).to(dtype)
    log_u = 0.53 + 0.74 * r(S, C, Z)
    obs = 1.2 + 0.5 * r(S, C, 1) + 0.3 * basement + 0.4 * log_u + 0.7 * r(S, C, Z)
    nm = lambda x, *names: x.refine_names(*names)
    params = {
        'gm_loc': 0.1 * r(), 'gm_ls': -0.3 + 0.1 * r(), 'gls_loc': 0.1 * r(), 'gls_ls': -0.3 + 0.1 * r(),
        'sm_loc': nm(1.0 + 0.1 * r(S), 'States'), 'sm_ls': nm(-0.5 + 0.1 * r(S), 'States'),
        'sls_loc': nm(-0.5 + 0.1 * r(S), 'States'), 'sls_ls': nm(-0.7 + 0.1 * r(S), 'States'),
        'cm_loc': nm(1.0 + 0.1 * r(S, C), 'States', 'Counties'), 'cm_ls': nm(-0.7 + 0.1 * r(S, C), 'States', 'Counties'),
        'cls_loc': nm(-0.3 + 0.1 * r(S, C), 'States', 'Counties'), 'cls_ls': nm(-1. + 0.1 * r(S, C), 'States', 'Counties'),
        'bu_loc': nm(0.3 + 0.1 * r(S, C), 'States', 'Counties'), 'bu_ls': nm(-1. + 0.1 * r(S, C), 'States', 'Counties'),
        'bb_loc': nm(0.3 + 0.1 * r(S, C), 'States', 'Counties'), 'bb_ls': nm(-1. + 0.1 * r(S, C), 'States', 'Counties'),
    }
    return dict(platesizes={'States': S, 'Counties': C, 'Zips': Z},
                data={'obs': nm(obs, 'States', 'Counties', 'Zips')},
                inputs={'basement': nm(basement, 'States', 'Counties', 'Zips'),
                        'log_uranium': nm(log_u, 'States', 'Counties', 'Zips')},
                params=params)


# --------------------------------------------------------------------------- cfg4
def timeseries_model(ns):
    """tests/timeseries.py:15-30 (A=0.9, noise 0.1, obs noise 1)."""
    P = ns.Plate(
        init=ns.Normal(0, 1.),
        T=ns.Plate(
            ts=ns.Timeseries("init", ns.Normal(lambda prev: 0.9 * prev, 0.1)),
            obs=ns.Normal('ts', 1.),
        ),
    )
    Q = ns.Plate(
        init=ns.Normal(0, 1),
        T=ns.Plate(
            ts=ns.Normal(0, 1),
            obs=ns.Data(),
        ),
    )
    return P, Q


def timeseries_inputs(T=1000, seed=0, dtype=t.float32):
    g = t.Generator().manual_seed(seed)
    x = t.zeros(T, dtype=t.float64)
    prev = t.randn((), generator=g, dtype=t.float64)
    for i in range(T):
        prev = 0.9 * prev + 0.1 * t.randn((), generator=g, dtype=t.float64)
        x[i] = prev
    obs = (x + t.randn(T, generator=g, dtype=t.float64)).to(dtype)
    return dict(platesizes={'T': T}, data={'obs': obs.refine_names('T')}, inputs={}, params={})


# --------------------------------------------------------------------------- extra structure
def model1_model(ns):
    """tests/model1.py:5-30: Group in a nested plate, latent-dependent scale, dangling latents."""
    P = ns.Plate(
        ab=ns.Group(
            a=ns.Normal(0, 1),
            b=ns.Normal("a", 1),
        ),
        c=ns.Normal(0, lambda a: a.exp()),
        p1=ns.Plate(
            d=ns.Normal("a", 1),
            p2=ns.Plate(
                e=ns.Normal("d", 1.),
            ),
        ),
    )
    Q = ns.Plate(
        ab=ns.Group(
            a=ns.Normal("a_mean", 1),
            b=ns.Normal("a", 1),
        ),
        c=ns.Normal(0, lambda a: a.exp()),
        p1=ns.Plate(
            d=ns.Normal("d_mean", 1),
            p2=ns.Plate(
                e=ns.Data(),
            ),
        ),
    )
    return P, Q


def model1_inputs(p1=3, p2=4, seed=0, dtype=t.float32):
    g = t.Generator().manual_seed(seed)
    r = lambda *s: t.randn(s, generator=g, dtype=t.float64).to(dtype)
    return dict(platesizes={'p1': p1, 'p2': p2},
                data={'e': r(p1, p2).refine_names('p1', 'p2')},
                inputs={},
                params={'a_mean': 0.1 * r(), 'd_mean': (0.1 * r(p1)).refine_names('p1')})


# --------------------------------------------------------------------------- more of the reference's own test models
def ref_bernoulli_model(ns):
    """tests/bernoulli_no_plate.py:5-17 (Beta prior, Bernoulli(probs) likelihood)."""
    P = ns.Plate(p=ns.Beta(2, 1), T=ns.Plate(coin=ns.Bernoulli('p')))
    Q = ns.Plate(p=ns.Beta(1, 1), T=ns.Plate(coin=ns.Data()))
    return P, Q


def ref_bernoulli_inputs(T=10, seed=0, dtype=t.float32):
    coin = t.cat([t.zeros(3), t.ones(T - 3)]).to(dtype)
    return dict(platesizes={'T': T}, data={'coin': coin.refine_names('T')}, inputs={}, params={})


def ref_corr_q_model(ns):
    """tests/linear_gaussian_two_params_corr_Q.py:31-46 (two unplated latents, correlated Q)."""
    P = ns.Plate(a=ns.Normal(2, 1), b=ns.Normal('a', 1), T=ns.Plate(d=ns.Normal('b', 3)))
    Q = ns.Plate(a=ns.Normal(1, 4), b=ns.Normal('a', 1.2), T=ns.Plate(d=ns.Data()))
    return P, Q


def ref_dangling_model(ns):
    """tests/linear_gaussian_latents_dangling.py:28-47 (a plated latent nothing depends on)."""
    P = ns.Plate(
        a=ns.Normal(2, 2),
        T=ns.Plate(z=ns.Normal('a', 1.3), zp=ns.Normal('a', 1.), d=ns.Normal('z', 1.5)),
    )
    Q = ns.Plate(
        a=ns.Normal(1, 4),
        T=ns.Plate(z=ns.Normal(lambda a: 1.5 * a, 3.5), zp=ns.Normal(lambda a: 1.5 * a, 3.5), d=ns.Data()),
    )
    return P, Q


def ref_batch_model(ns):
    """tests/linear_gaussian_latents_batch.py:22-37 (event shape [2], tensor-valued constant arguments)."""
    P = ns.Plate(
        a=ns.Normal(t.tensor([0.3, -0.8]), t.tensor([1., 2.])),
        T=ns.Plate(z=ns.Normal('a', t.tensor([1.3, 1.6])), d=ns.Normal('z', t.tensor([2., 3.]))),
    )
    Q = ns.Plate(
        a=ns.Normal(t.zeros(2), 4),
        T=ns.Plate(z=ns.Normal(lambda a: 0.5 * a, 6), d=ns.Data()),
    )
    return P, Q


def ref_batch_inputs(T=10, seed=0, dtype=t.float32):
    g = t.Generator().manual_seed(seed)
    data = {'d': (1.5 + t.randn(T, 2, generator=g, dtype=t.float64)).to(dtype).refine_names('T', None)}
    return dict(platesizes={'T': T}, data=data, inputs={}, params={})


CASES = {
    # name: (model builder, inputs builder, kwargs for inputs, K, moment specs, joints, importance N or None)
    'cfg1_lgl': (lgl_model, lgl_inputs, dict(T=10), 3, [('a', 'mean'), ('a', 'mean2'), ('z', 'mean'), ('z', 'mean2')], [], 7),
    'cfg1_lglp': (lglp_model, lglp_inputs, dict(T=10), 4, [('a', 'mean'), ('z', 'mean2')], [], 5),
    'cfg2_movielens': (movielens_model, movielens_inputs, dict(M=12, N=5, d=18), 6,
                       [('z', 'mean'), ('mu_z', 'mean2')], [('mu_z', 'psi_z')], 6),
    'cfg3_radon': (radon_model, radon_inputs, dict(S=3, C=4, Z=5), 4,
                   [('County_mean', 'mean'), ('global_mean', 'mean2')], [('Beta_u', 'Beta_basement')], 9),
    'cfg4_timeseries': (timeseries_model, timeseries_inputs, dict(T=37), 5, [('ts', 'mean'), ('ts', 'mean2')], [], None),
    'model1': (model1_model, model1_inputs, dict(p1=3, p2=4), 4, [('d', 'mean'), ('c', 'mean2')], [('ab', 'c')], 5),
    'ref_bernoulli': (ref_bernoulli_model, ref_bernoulli_inputs, dict(T=10), 5, [('p', 'mean')], [], 6),
    'ref_corr_q': (ref_corr_q_model, lgl_inputs, dict(T=10), 4, [('a', 'mean'), ('b', 'mean2')], [('a', 'b')], 5),
    'ref_dangling': (ref_dangling_model, lgl_inputs, dict(T=10), 3, [('a', 'mean'), ('zp', 'mean')], [], 4),
    'ref_batch': (ref_batch_model, ref_batch_inputs, dict(T=10), 3, [('a', 'mean'), ('z', 'mean2')], [], 4),
}

MOMENT_FUNCS = {'mean': (lambda x: x), 'mean2': (lambda x: x * x)}


def build(case, ns, dtype=t.float32):
    """Build the (P, Q) of a case with tensor-valued constants created in `dtype` (the golden generator builds the
    reference model under torch.set_default_dtype(dtype), so `t.tensor([1.3, 1.6])` means different bits per tag)."""
    old = t.get_default_dtype()
    t.set_default_dtype(dtype)
    try:
        return CASES[case][0](ns)
    finally:
        t.set_default_dtype(old)


# --------------------------------------------------------------------------- QEM
def qem_model1_model(ns):
    """/root/reference/tests/model1.py:5-30 with its own parameter declarations: QEMParam on a Group member,
    OptParam and a named extra parameter in a plate."""
    P = ns.Plate(
        ab=ns.Group(
            a=ns.Normal(0, 1),
            b=ns.Normal("a", 1),
        ),
        c=ns.Normal(0, lambda a: a.exp()),
        p1=ns.Plate(
            d=ns.Normal("a", 1),
            p2=ns.Plate(
                e=ns.Normal("d", 1.),
            ),
        ),
    )
    Q = ns.Plate(
        ab=ns.Group(
            a=ns.Normal(ns.QEMParam(0.), ns.QEMParam(1.)),
            b=ns.Normal("a", 1),
        ),
        c=ns.Normal(0, lambda a: a.exp()),
        p1=ns.Plate(
            d=ns.Normal(ns.OptParam(0.), "d_scale"),
            p2=ns.Plate(
                e=ns.Data(),
            ),
        ),
    )
    return P, Q


def qem_model1_inputs(p1=4, p2=4, seed=0, dtype=t.float32):
    g = t.Generator().manual_seed(seed)
    r = lambda *s: t.randn(s, generator=g, dtype=t.float64).to(dtype)
    return dict(platesizes={'p1': p1, 'p2': p2}, data={'e': r(p1, p2).refine_names('p1', 'p2')}, inputs={},
                params={'d_scale': t.ones(p1, dtype=dtype).refine_names('p1')})


def qem_families_model(ns):
    """Every family with a mean <-> conventional conversion on the device (conversions.py:46-296): global and plated
    QEM latents, a QEM distribution on the P side too (BoundPlate._update_qem_params runs for P and for Q)."""
    P = ns.Plate(
        g=ns.Gamma(2., 2.),
        bt=ns.Beta(2., 3.),
        ex=ns.Exponential(1.5),
        hn=ns.HalfNormal(1.2),
        m=ns.Normal(ns.QEMParam(0.3), ns.QEMParam(1.5)),
        p1=ns.Plate(
            w=ns.Gamma(3., 3.),
            bn=ns.Bernoulli(probs=0.3),
            v=ns.Normal('m', 1.),
            y=ns.Normal(lambda g, ex, w, bn, bt, v: g - ex + w * bn + bt + v, lambda hn: hn + 0.5),
        ),
    )
    Q = ns.Plate(
        g=ns.Gamma(ns.QEMParam(2.5), ns.QEMParam(2.)),
        bt=ns.Beta(ns.QEMParam(2.), ns.QEMParam(2.)),
        ex=ns.Exponential(ns.QEMParam(1.)),
        hn=ns.HalfNormal(ns.QEMParam(1.)),
        m=ns.Normal(0., 2.),
        p1=ns.Plate(
            w=ns.Gamma(ns.QEMParam(3.), ns.QEMParam(2.5)),
            bn=ns.Bernoulli(probs=ns.QEMParam(0.4)),
            v=ns.Normal(ns.QEMParam(0.), ns.QEMParam(1.)),
            y=ns.Data(),
        ),
    )
    return P, Q


def qem_families_inputs(p1=6, seed=0, dtype=t.float32):
    g = t.Generator().manual_seed(seed)
    return dict(platesizes={'p1': p1}, data={'y': (0.5 + t.randn(p1, generator=g, dtype=t.float64)).to(dtype).refine_names('p1')},
                inputs={}, params={})


QEM_CASES = {
    # name: (model builder, inputs builder, kwargs, K, learning rates of the successive updates)
    'qem_model1': (qem_model1_model, qem_model1_inputs, dict(p1=4, p2=4), 10, (0.1, 0.3)),
    'qem_families': (qem_families_model, qem_families_inputs, dict(p1=6), 12, (0.2, 0.5, 0.1)),
}


# --------------------------------------------------------------------------- MultivariateNormal (row f-2)
def _mvn_consts(F, dtype):
    g = t.Generator().manual_seed(1234 + F)
    r = lambda *s: t.randn(s, generator=g, dtype=t.float64)
    A, B, C = r(F, F), r(F, F), r(F, F)
    return dict(prior_mean=r(F).to(dtype), prior_cov=(A @ A.mT + 0.5 * t.eye(F, dtype=t.float64)).to(dtype),
                ap_mean=r(F).to(dtype), ap_cov=(B @ B.mT + 2 * t.eye(F, dtype=t.float64)).to(dtype),
                like_cov=(C @ C.mT + 0.5 * t.eye(F, dtype=t.float64)).to(dtype))


def mvn_model(ns, F=3):
    """/root/reference/tests/linear_multivariate_gaussian_param.py:25-45: a MultivariateNormal latent with a plated
    MultivariateNormal likelihood; constants positional (loc, covariance_matrix)."""
    c = _mvn_consts(F, t.get_default_dtype())
    P = ns.Plate(
        a=ns.MultivariateNormal(c['prior_mean'], c['prior_cov']),
        T=ns.Plate(
            d=ns.MultivariateNormal('a', c['like_cov']),
        ),
    )
    Q = ns.Plate(
        a=ns.MultivariateNormal('qa_loc', c['ap_cov']),
        T=ns.Plate(
            d=ns.Data(),
        ),
    )
    return P, Q


def mvn_inputs(T=10, F=3, seed=0, dtype=t.float32):
    g = t.Generator().manual_seed(seed)
    r = lambda *s: t.randn(s, generator=g, dtype=t.float64).to(dtype)
    return dict(platesizes={'T': T}, data={'d': (1.5 + r(T, F)).refine_names('T', None)}, inputs={},
                params={'qa_loc': _mvn_consts(F, dtype)['ap_mean'].clone()})


def mvn2_model(ns, F=4):
    """precision_matrix and scale_tril arguments, a latent-dependent loc through a lambda, a plated MvN latent whose
    Q depends on its parent (mixture Q), Normal data on one coordinate pair."""
    c = _mvn_consts(F, t.get_default_dtype())
    prec = t.linalg.inv(c['prior_cov'].double()).to(c['prior_cov'].dtype)
    prec = 0.5 * (prec + prec.mT)
    tril = t.linalg.cholesky(c['like_cov'].double()).to(c['like_cov'].dtype)
    P = ns.Plate(
        a=ns.MultivariateNormal(c['prior_mean'], precision_matrix=prec),
        T=ns.Plate(
            z=ns.MultivariateNormal(lambda a: 0.5 * a, scale_tril=tril),
            d=ns.MultivariateNormal('z', c['ap_cov']),
        ),
    )
    Q = ns.Plate(
        a=ns.MultivariateNormal('qa_loc', scale_tril=tril),
        T=ns.Plate(
            z=ns.MultivariateNormal(lambda a, qz_b: 0.3 * a + qz_b, precision_matrix=prec),
            d=ns.Data(),
        ),
    )
    return P, Q


def mvn2_inputs(T=6, F=4, seed=0, dtype=t.float32):
    g = t.Generator().manual_seed(seed)
    r = lambda *s: t.randn(s, generator=g, dtype=t.float64).to(dtype)
    return dict(platesizes={'T': T}, data={'d': (0.5 + r(T, F)).refine_names('T', None)}, inputs={},
                params={'qa_loc': _mvn_consts(F, dtype)['ap_mean'].clone(), 'qz_b': (0.1 * r(T, F)).refine_names('T', None)})


CASES['mvn'] = (mvn_model, mvn_inputs, dict(T=10), 5, [('a', 'mean'), ('a', 'mean2')], [], 6)
CASES['mvn2'] = (mvn2_model, mvn2_inputs, dict(T=6), 4, [('a', 'mean'), ('z', 'mean2')], [], 5)


# --------------------------------------------------------------------------- Dirichlet (row f-2)
def dirichlet_model(ns):
    """A simplex-valued global latent and a plated one whose concentration depends on it (dist.py:323-359 family list)."""
    P = ns.Plate(
        p=ns.Dirichlet(t.tensor([1.5, 2.0, 0.7, 3.0])),
        T=ns.Plate(
            q=ns.Dirichlet(lambda p: 1.0 + 4.0 * p),
            y=ns.Normal(lambda q, w: q @ w, 0.7),
        ),
    )
    Q = ns.Plate(
        p=ns.Dirichlet('qp_conc'),
        T=ns.Plate(
            q=ns.Dirichlet(lambda p, qq_w: 0.5 + qq_w * p),
            y=ns.Data(),
        ),
    )
    return P, Q


def dirichlet_inputs(T=5, seed=0, dtype=t.float32):
    g = t.Generator().manual_seed(seed)
    return dict(platesizes={'T': T}, data={'y': t.randn(T, generator=g, dtype=t.float64).to(dtype).refine_names('T')},
                inputs={'w': t.tensor([1.0, -2.0, 0.5, 3.0], dtype=dtype)},
                params={'qp_conc': t.tensor([2.0, 2.5, 1.2, 2.2], dtype=dtype), 'qq_w': t.tensor(3.0, dtype=dtype)})


CASES['dirichlet'] = (dirichlet_model, dirichlet_inputs, dict(T=5), 4, [('p', 'mean'), ('q', 'mean2')], [], 5)


# --------------------------------------------------------------------------- vector-valued likelihoods (row f-2)
def vector_family_case(ns, family, dtype, seed=0, T=9, K=5, J=4):
    """A simplex-valued latent `p` and a vector of scores `s` feeding a OneHotCategorical / Multinomial / Categorical likelihood
    (through probs and through logits): returns (P, Q, sample, params, data) as named tensors."""
    from alan_b200.named import NT
    Fam = getattr(ns, family)
    if family != 'RelaxedOneHotCategorical':
        P = ns.Plate(p=ns.Dirichlet(t.tensor([1.5, 2.0, 0.7, 3.0], dtype=dtype)), s=ns.Normal(t.zeros(J, dtype=dtype), 1.0),
                     T=ns.Plate(y=Fam(probs='p'), w=Fam(logits=lambda s: 1.5 * s)))
    Q = ns.Plate(p=ns.Dirichlet('qp_conc'), s=ns.Normal('s_loc', lambda s_ls: s_ls.exp()),
                 T=ns.Plate(y=ns.Data(), w=ns.Data()))
    g = t.Generator().manual_seed(seed)
    r = lambda *sh: t.randn(sh, generator=g, dtype=t.float64).to(dtype)
    if family == 'RelaxedOneHotCategorical':                      # points of the open simplex, temperature first
        P = ns.Plate(p=ns.Dirichlet(t.tensor([1.5, 2.0, 0.7, 3.0], dtype=dtype)), s=ns.Normal(t.zeros(J, dtype=dtype), 1.0),
                     T=ns.Plate(y=Fam(0.7, probs='p'), w=Fam(1.3, logits=lambda s: 1.5 * s)))
        dir_ = t.distributions.Dirichlet(t.full((J,), 2.0, dtype=t.float64))
        onehot = lambda: dir_.sample((T,)).to(dtype)
    elif family == 'Categorical':                                 # the value is the class index itself
        onehot = lambda: t.randint(0, J, (T,), generator=g).to(dtype)
    else:
        onehot = lambda: t.nn.functional.one_hot(t.randint(0, J, (T,), generator=g), J).to(dtype)
    data = {'y': NT(onehot(), ('T',)), 'w': NT(onehot(), ('T',))}
    params = {'qp_conc': NT(t.tensor([2.0, 2.5, 1.2, 2.2], dtype=dtype), ()), 's_loc': NT(0.1 * r(J), ()),
              's_ls': NT(-0.5 + 0.1 * r(J), ())}
    p = t.distributions.Dirichlet(t.tensor([2.0, 2.5, 1.2, 2.2], dtype=t.float64)).sample((K,)).to(dtype)
    sample = {'p': NT(p, ('K_p',)), 's': NT(0.6 * r(K, J), ('K_s',))}
    return P, Q, sample, params, data


# --------------------------------------------------------------------------- Categorical + composed families (row f-2)
def families_model(ns, J=4, F=3):
    """A softmax-regression likelihood like the reference's examples/examples/mnist_classification.py:30-40
    (`Categorical(logits=...)` of a matrix-valued latent times a plated covariate), next to Gumbel and Weibull
    likelihoods of a scalar latent: families whose density the B200 planner composes from VM primitives."""
    dt = t.get_default_dtype()
    P = ns.Plate(
        w=ns.Normal(t.zeros(J, F, dtype=dt), 1.0),
        s=ns.Normal(0., 1.),
        T=ns.Plate(
            y=ns.Categorical(logits=lambda w, x: w @ x),
            g=ns.Gumbel('s', 1.3),
            wb=ns.Weibull(lambda s: s.exp(), 1.7),
        ),
    )
    Q = ns.Plate(
        w=ns.Normal('w_loc', lambda w_ls: w_ls.exp()),
        s=ns.Normal('s_loc', 0.8),
        T=ns.Plate(y=ns.Data(), g=ns.Data(), wb=ns.Data()),
    )
    return P, Q


def families_inputs(T=7, J=4, F=3, seed=0, dtype=t.float32):
    g = t.Generator().manual_seed(seed)
    r = lambda *s: t.randn(s, generator=g, dtype=t.float64).to(dtype)
    return dict(platesizes={'T': T},
                data={'y': t.randint(0, J, (T,), generator=g).to(dtype).refine_names('T'),
                      'g': (0.3 + 1.2 * r(T)).refine_names('T'), 'wb': (r(T).abs() + 0.2).refine_names('T')},
                inputs={'x': r(T, F).refine_names('T', None)},
                params={'w_loc': 0.2 * r(J, F), 'w_ls': -0.6 + 0.1 * r(J, F), 's_loc': 0.1 * r()})


CASES['families'] = (families_model, families_inputs, dict(T=7), 4, [('s', 'mean'), ('w', 'mean2')], [], 5)


# --------------------------------------------------------------------------- LowRankMultivariateNormal (row f-2)
def _lowrank_consts(F, R, dtype):
    g = t.Generator().manual_seed(4321 + F)
    r = lambda *s: t.randn(s, generator=g, dtype=t.float64)
    return dict(prior_mean=r(F).to(dtype), prior_fac=(0.8 * r(F, R)).to(dtype), prior_diag=(0.5 + r(F).abs()).to(dtype),
                ap_mean=r(F).to(dtype), ap_fac=(0.6 * r(F, R)).to(dtype), ap_diag=(1.0 + r(F).abs()).to(dtype),
                like_fac=(0.7 * r(F, R)).to(dtype), like_diag=(0.4 + r(F).abs()).to(dtype))


def lowrank_model(ns, F=4, R=2):
    """The reference's multivariate-Gaussian test model (tests/linear_multivariate_gaussian_param.py:25-45) with every
    covariance given in low-rank-plus-diagonal form (dist.py:323-359 `LowRankMultivariateNormal`, positional arguments
    loc, cov_factor, cov_diag)."""
    c = _lowrank_consts(F, R, t.get_default_dtype())
    P = ns.Plate(
        a=ns.LowRankMultivariateNormal(c['prior_mean'], c['prior_fac'], c['prior_diag']),
        T=ns.Plate(d=ns.LowRankMultivariateNormal('a', c['like_fac'], c['like_diag'])),
    )
    Q = ns.Plate(
        a=ns.LowRankMultivariateNormal('qa_loc', c['ap_fac'], c['ap_diag']),
        T=ns.Plate(d=ns.Data()),
    )
    return P, Q


def lowrank_inputs(T=8, F=4, R=2, seed=0, dtype=t.float32):
    g = t.Generator().manual_seed(seed)
    r = lambda *s: t.randn(s, generator=g, dtype=t.float64).to(dtype)
    return dict(platesizes={'T': T}, data={'d': (0.5 + r(T, F)).refine_names('T', None)}, inputs={},
                params={'qa_loc': _lowrank_consts(F, R, dtype)['ap_mean'].clone()})


CASES['lowrank'] = (lowrank_model, lowrank_inputs, dict(T=8), 5, [('a', 'mean'), ('a', 'mean2')], [], 6)
