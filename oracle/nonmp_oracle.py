"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the reference's global importance sampling (`SampleNonMP`).

Plain PyTorch restatement, function by function (paths relative to /root/reference/):
  unify_K               src/alan/SampleNonMP.py:127-137  (`unify_dims`: every K axis becomes the one axis K)
  non_mp_log_prob       src/alan/SampleNonMP.py:139-207  (log P - log Q per variable, summed over the plates)
  elbo                  src/alan/SampleNonMP.py:56-57    (`logsumexp` over K WITHOUT eps, minus log K)
  weights / moments     src/alan/SampleNonMP.py:100-116  (softmax over K; RawMoment.from_marginals, moments.py:16-35)
  importance_sample_idxs src/alan/SampleNonMP.py:71-90   (categorical over K with weights exp(lpq - max))

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file; the product never does.
Parity pin: tests/golden/nonmp_*.pt hold outputs of the UNMODIFIED reference (tests/golden/make_golden_nonmp.py);
tests/test_nonmp_cpu.py checks this file against every one of them.

Deviation (as in oracle/logpq_oracle.py): the categorical draw takes EXPLICIT float64 uniforms and the inverse-CDF rule
of SURVEY.md Appendix A8 instead of torch.multinomial's RNG stream.
"""
from __future__ import annotations

import math

import torch as t

from alan_b200.model import Plate, Data, Timeseries
from alan_b200.named import NT
from .logpq_oracle import ONT, ont, dist_log_prob, inverse_cdf_draw, _working_dtype, _cast

K_AXIS = 'K'


def unify_K(sample: dict) -> dict:
    out = {}
    for k, v in sample.items():
        ks = [a for a in v.axes if a.startswith('K_')]
        assert len(ks) == 1, f"{k}: exactly one K axis expected, got {ks}"
        out[k] = ONT(v.t, tuple(K_AXIS if a == ks[0] else a for a in v.axes))
    return out


def non_mp_log_prob(P: Plate, Q: Plate, sample: dict, scope: dict, data: dict, active: tuple, dtype,
                    extra: dict | None = None) -> ONT:
    """SampleNonMP.py:139-207.  `sample` / `data` are flat {name: ONT}; returns an ONT with axes (K,)."""
    scope = dict(scope)
    for k in Q.flat_prog:                                   # update_scope(scope, sample): this level's samples
        if k in sample:
            scope[k] = sample[k]
    total = None

    def add(x):
        nonlocal total
        total = x if total is None else total + x

    for key, e in (extra or {}).items():                    # zero source terms f(x) * J of the moments, at their level
        plates = set(a for a in e.axes if a != K_AXIS)
        if plates == set(active):
            add(e.sum_pos().sum(tuple(active)) if active else e.sum_pos())
    for k, dQ in Q.flat_prog.items():
        dP = P.flat_prog[k]
        assert not isinstance(dP, Timeseries)
        if isinstance(dQ, Plate):
            lpq = non_mp_log_prob(dP, dQ, sample, scope, data, (*active, k), dtype, extra)
        elif isinstance(dQ, Data):
            lpq = dist_log_prob(dP, data[k], scope, dtype)
            assert set(lpq.axes) == set(active) | {K_AXIS}
            lpq = lpq.sum(tuple(active)) if active else lpq
        else:
            lp = dist_log_prob(dP, sample[k], scope, dtype)
            lq = dist_log_prob(dQ, sample[k], scope, dtype)
            assert set(lp.axes) == set(active) | {K_AXIS} and set(lq.axes) == set(active) | {K_AXIS}
            if active:
                lp, lq = lp.sum(tuple(active)), lq.sum(tuple(active))
            lpq = lp - lq
        add(lpq)
    assert tuple(total.axes) == (K_AXIS,)
    return total


def logpq(P, Q, sample, inputs_params, data, extra=None) -> ONT:
    s = unify_K(sample)
    dtype = _working_dtype(s, {k: ont(v) for k, v in (inputs_params or {}).items()},
                           {k: ont(v) for k, v in (data or {}).items()})
    s = _cast(s, dtype)
    ip = _cast({k: ont(v) for k, v in (inputs_params or {}).items()}, dtype)
    d = _cast({k: ont(v) for k, v in (data or {}).items()}, dtype)
    return non_mp_log_prob(P, Q, s, ip, d, (), dtype, extra)


def elbo(P, Q, sample, inputs_params, data, extra=None) -> t.Tensor:
    lpq = logpq(P, Q, sample, inputs_params, data, extra)
    return t.logsumexp(lpq.order((K_AXIS,)).t, 0) - math.log(lpq.sizes()[K_AXIS])


def weights(P, Q, sample, inputs_params, data) -> t.Tensor:
    """(lpq - lpq.logsumexp(K)).exp()  (SampleNonMP.py:106-107)."""
    lpq = logpq(P, Q, sample, inputs_params, data).order((K_AXIS,)).t
    return (lpq - t.logsumexp(lpq, 0)).exp()


def moments(P, Q, sample, inputs_params, data, moms):
    """moms: [(varnames tuple, f)] -> list of NT `sum_K w_k f(x_k)` with axes = plates (moments.py:16-35)."""
    w = weights(P, Q, {k: v.detach() for k, v in sample.items()}, inputs_params, data)
    s = unify_K({k: v.detach() for k, v in sample.items()})
    all_plates = P.all_platenames()
    out = []
    for varnames, f in moms:
        fx = f(*[ONT(s[v].t.to(w.dtype), s[v].axes) for v in varnames])
        plates = tuple(a for a in all_plates if a in fx.axes)
        fx = fx.order((K_AXIS,) + plates)
        wv = w.reshape([-1] + [1] * (fx.t.ndim - 1))
        out.append(NT((fx.t * wv).sum(0), plates))
    return out


def importance_sample_idxs(P, Q, sample, inputs_params, data, u: t.Tensor) -> t.Tensor:
    """u: float64 uniforms [N] -> int64 indices [N] over the K axis."""
    lpq = logpq(P, Q, {k: v.detach() for k, v in sample.items()}, inputs_params, data)
    return inverse_cdf_draw(lpq, (K_AXIS,), ONT(u.to(t.float64), ('N',)))[K_AXIS].order(('N',)).t
