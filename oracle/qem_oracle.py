"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the reference's QEM update.

Plain PyTorch restatement (paths relative to /root/reference/):
  inverse_digamma, *Conversion.mean2conv / conv2mean     src/alan/conversions.py:8-35, 46-296
  update_moving_avg                                      src/alan/BoundPlate.py:272-286
  update_convparams                                      src/alan/BoundPlate.py:256-270
  update                                                 src/alan/Sample.py:351-355 (P first, then Q)
The posterior moments come from oracle/logpq_oracle.py `moments` (Sample.py:291-346).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file; the product never does.
Parity pin: tests/golden/qem_*.pt hold the QEM parameters and means of the UNMODIFIED reference after successive
`sample.update_qem_params(lr)` calls (tests/golden/make_golden_qem.py); tests/test_qem_cpu.py checks this file
against them.
"""
from __future__ import annotations

import torch as t

from alan_b200.named import NT
from . import logpq_oracle as O

MOMENT_FUNCS = {
    'mean': lambda x: x,
    'mean2': lambda x: x * x,
    'mean_log': lambda x: x.log(),
    'mean_log1m': lambda x: (1 - x).log(),
}
SUFFICIENT = {
    'Normal': ('mean', 'mean2'), 'Bernoulli': ('mean',), 'Poisson': ('mean',), 'Exponential': ('mean',),
    'HalfNormal': ('mean2',), 'Gamma': ('mean_log', 'mean'), 'Beta': ('mean_log', 'mean_log1m'),
}


def grad_digamma(x):
    return t.special.polygamma(1, x)


def inverse_digamma(y):
    """conversions.py:8-35 (Minka, appendix C)."""
    x = t.where(y > -2.22, y.exp() + 0.5, -t.reciprocal(y - t.digamma(t.ones((), dtype=y.dtype))))
    for _ in range(6):
        x = x - (t.digamma(x) - y) / grad_digamma(x)
    return x


def dirichlet_mean2conv(logp):
    """conversions.py:127-147."""
    alpha = t.ones_like(logp)
    for _ in range(5):
        alpha = inverse_digamma(t.digamma(alpha.sum(-1, keepdim=True)) + logp)
    for _ in range(6):
        sum_alpha = alpha.sum(-1, keepdim=True)
        g = t.digamma(sum_alpha) - t.digamma(alpha) + logp
        z = grad_digamma(sum_alpha)
        q = -grad_digamma(alpha)
        b = (g / q).sum(-1, keepdim=True) / (1 / z + (1 / q).sum(-1, keepdim=True))
        alpha = alpha - (g - b) / q
    return alpha


def mean2conv(family, means):
    if family == 'Normal':                                           # :93-98
        mean, mean2 = means
        return {'loc': mean, 'scale': (mean2 - mean * mean).sqrt().clamp(min=t.finfo(mean2.dtype).tiny)}
    if family == 'Bernoulli':                                        # :60-62
        return {'probs': means[0]}
    if family == 'Poisson':                                          # :79-81
        return {'rate': means[0]}
    if family == 'Exponential':                                      # :109-111
        return {'rate': t.reciprocal(means[0])}
    if family == 'HalfNormal':                                       # :288-290
        return {'scale': means[0].sqrt()}
    if family == 'Gamma':                                            # :204-218
        Elogx, Ex = means
        diff = Elogx - Ex.log()
        alpha = -0.5 / diff
        for _ in range(6):
            num = diff + alpha.log() - t.digamma(alpha)
            denom = 1 - alpha * grad_digamma(alpha)
            alpha = alpha * t.reciprocal(1 + num / denom)
        return {'concentration': alpha, 'rate': alpha / Ex}
    if family == 'Beta':                                             # :169-175
        c = dirichlet_mean2conv(t.stack([means[0], means[1]], -1))
        return {'concentration1': c[..., 0], 'concentration0': c[..., 1]}
    raise Exception(family)


def conv2mean(family, a):
    if family == 'Normal':
        return (a['loc'], a['loc'] ** 2 + a['scale'] ** 2)
    if family in ('Bernoulli',):
        return (a['probs'],)
    if family == 'Poisson':
        return (a['rate'],)
    if family == 'Exponential':
        return (t.reciprocal(a['rate']),)
    if family == 'HalfNormal':
        return (a['scale'] ** 2,)
    if family == 'Gamma':
        return (-t.log(a['rate']) + t.digamma(a['concentration']), a['concentration'] / a['rate'])
    if family == 'Beta':
        norm = t.digamma(a['concentration1'] + a['concentration0'])
        return (t.digamma(a['concentration1']) - norm, t.digamma(a['concentration0']) - norm)
    raise Exception(family)


def update_side(qvars, params: dict, means: dict, lr, P, Q, sample, inputs_params, data):
    """One side's `_update_qem_params` (BoundPlate.py:288-290).  qvars: [(varname, family, {argname: paramname},
    (meannames...))]; params / means: {name: NT}, updated IN PLACE like the reference's buffers."""
    if not qvars:
        return
    moms = [((v,), MOMENT_FUNCS[s]) for v, fam, _, _ in qvars for s in SUFFICIENT[fam]]
    new = O.moments(P, Q, sample, inputs_params, data, moms)
    i = 0
    for v, fam, arg2param, meannames in qvars:
        ms = []
        for mn in meannames:
            prev = means[mn]
            prev.t.mul_(1 - lr).add_(new[i].order(prev.axes).t.to(prev.t.dtype), alpha=lr)
            ms.append(prev.t)
            i += 1
        for arg, val in mean2conv(fam, ms).items():
            params[arg2param[arg]].t.copy_(val)
