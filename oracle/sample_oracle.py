"""TEST INFRASTRUCTURE ONLY -- CPU oracle for ancestral sampling of Q.

Plain-torch restatement of the reference's sampling walk with EXPLICIT base noise:
  sample_plate     Plate.sample                          src/alan/Plate.py:93-143
  sample_gdt       sample_gdt                            src/alan/dist.py:23-72
  resample_scope   Sampler.resample_scope / .perm        src/alan/Sampler.py:85-116,139-160
  draw             Dist.sample -> TorchDimDist.sample    src/alan/dist.py:304-309, TorchDimDist.py:88-125
  timeseries_draw  Timeseries.sample                     src/alan/Timeseries.py:89-123

Noise contract (the same keys alan_b200.sampling.QSampler.noise_shapes() lists):
  noise[(group, parent K axis)]  float64 uniforms [parent plates..., K]: the permutation of that parent's particles
                                 = argsort along K (PermutationSampler) / floor(u K) (CategoricalSampler)
  noise[(group, 'timeseries')]   float64 uniforms [active plates..., K]: timeseries_perm
  noise[varname]                 standard normals / uniforms [active plates..., K, *event]: the draw's base noise;
                                 draw = the closed-form transform below (Normal: loc + scale * eps, the formula of
                                 torch's rsample; the others by inverse CDF)

Parity pin: tests/golden/qsample_*.pt hold samples produced by the UNMODIFIED reference walk
(`Problem.Q._sample`) with `PermutationSampler.perm` and `TorchDimDist.sample` replaced by these same explicit-noise
primitives (tests/golden/make_golden_sampling.py); tests/test_sampling_cpu.py checks this file against them.
"""
from __future__ import annotations

import torch as t

from alan_b200.model import Plate, Timeseries, datagroup, Kname
from alan_b200.named import NT
from .logpq_oracle import ONT, ont, resolve_arg, _align


def perm_from_uniform(u: t.Tensor, mode: int) -> t.Tensor:
    """[..., K] float64 uniforms -> [..., K] int64: argsort along K (ties by index) or floor(u K)."""
    K = u.shape[-1]
    if mode == 1:
        return (u * K).long().clamp(max=K - 1)
    return t.argsort(u, dim=-1, stable=True)


def transform(family, args: dict, noise):
    """The draw as a function of the (broadcast) distribution arguments and the base noise."""
    if family == 'Normal':
        return args['loc'] + args['scale'] * noise
    if family == 'LogNormal':
        return (args['loc'] + args['scale'] * noise).exp()
    if family == 'HalfNormal':
        return noise.abs() * args['scale']
    if family == 'Exponential':
        return -(-noise).log1p() / args['rate']
    if family == 'Uniform':
        return args['low'] + (args['high'] - args['low']) * noise
    if family == 'Laplace':
        s = noise - 0.5
        return args['loc'] - args['scale'] * (s / s.abs()) * (-(2.0 * s.abs())).log1p()
    if family == 'Bernoulli':
        p = args['probs'] if 'probs' in args else args['logits'].sigmoid()
        return (noise < p).to(noise.t.dtype if isinstance(noise, ONT) else noise.dtype)
    raise Exception(f"oracle: no transform for {family}")


def scale_tril_of(name, M: t.Tensor) -> t.Tensor:
    """torch.distributions.MultivariateNormal.__init__: the scale_tril it keeps for rsample."""
    if name == 'scale_tril':
        return M
    if name == 'covariance_matrix':
        return t.linalg.cholesky(M)
    Lf = t.linalg.cholesky(t.flip(M, (-2, -1)))                       # _precision_to_scale_tril
    L_inv = t.transpose(t.flip(Lf, (-2, -1)), -2, -1)
    return t.linalg.solve_triangular(L_inv, t.eye(M.shape[-1], dtype=M.dtype), upper=False)


def draw(dist, scope, eps: ONT, dtype) -> ONT:
    args = {k: resolve_arg(v, scope, dtype) for k, v in dist.args.items()}
    if dist.family == 'MultivariateNormal':                            # loc + scale_tril @ eps (torch's rsample)
        mname = [k for k in args if k != 'loc'][0]
        L = ONT(scale_tril_of(mname, args[mname].t), args[mname].axes)
        res = args['loc'] + (L @ eps)
        return ONT(res.order(eps.axes).t.expand(eps.t.shape).contiguous(), eps.axes)
    names = list(args)
    raw, axes = _align([eps] + [args[k] for k in names])
    out = transform(dist.family, dict(zip(names, raw[1:])), raw[0])
    res = ONT(out, axes)
    # result carries exactly the noise's named axes, in its order
    return ONT(res.order(eps.axes).t.expand(eps.t.shape).contiguous(), eps.axes)


def resample_scope(scope: dict, group, K_axis, noise, mode):
    """Sampler.resample_scope: every parent permuted along ITS K axis (per cell of its own plates), then renamed."""
    out, by_K = {}, {}
    for name, x in scope.items():
        ks = [a for a in x.axes if a.startswith('K_')]
        assert len(ks) <= 1
        by_K.setdefault(ks[0] if ks else None, []).append(name)
    for Kp, names in by_K.items():
        if Kp is None:
            for n in names:
                out[n] = scope[n]
            continue
        x0 = scope[names[0]]
        plates0 = tuple(a for a in x0.axes if a != Kp)
        if mode == 2:                                                       # IndependentSampler: arange (Sampler.py:162-169)
            Kn = x0.sizes()[Kp]
            perm = t.arange(Kn).expand([x0.sizes()[a] for a in plates0] + [Kn])
        else:
            perm = perm_from_uniform(noise[(group, Kp)], mode)              # [plates0..., K]
        for n in names:
            x = scope[n].order(plates0 + (Kp,))
            idx = perm.reshape(list(perm.shape) + [1] * x.pos_ndim).expand(list(perm.shape) + list(x.t.shape[len(x.axes):]))
            out[n] = ONT(t.gather(x.t, len(plates0), idx), plates0 + (K_axis,))
    return out


def timeseries_draw(ts: Timeseries, var, scope, active, K_axis, eps: ONT, ts_perm, dtype) -> ONT:
    T_axis = active[-1]
    other = tuple(active[:-1])
    prev = scope[ts.init]
    if set(prev.axes) != set(other) | {K_axis}:
        raise Exception(f"Initial state, {ts.init}, doesn't have the right dimensions for timeseries {var}")
    prev = prev.order(other + (K_axis,))
    eps_t = eps.order((T_axis,) + other + (K_axis,))
    T = eps_t.t.shape[0]
    steps = []
    for time in range(T):
        sc = {}
        for k, v in scope.items():
            if T_axis in v.axes:
                vo = v.order((T_axis,))
                v = ONT(vo.t[time], vo.axes[1:])
            sc[k] = v
        sc['prev'] = prev
        x = draw(ts.trans, sc, ONT(eps_t.t[time], other + (K_axis,)), dtype)
        steps.append(x.t)
        if ts_perm is not None:                                            # [active..., K] -> this step's [other..., K]
            p = ts_perm.movedim(len(other), 0)[time]
            idx = p.reshape(list(p.shape) + [1] * x.pos_ndim).expand(list(p.shape) + list(x.t.shape[len(x.axes):]))
            prev = ONT(t.gather(x.t, len(other), idx), x.axes)
        else:
            prev = x
    out = t.stack(steps, 0)                                                # [T, other..., K, event]
    return ONT(out, (T_axis,) + other + (K_axis,)).order(tuple(active) + (K_axis,))


def sample_plate(Q: Plate, active, scope, noise, K, mode, dtype, result):
    scope = dict(scope)
    for name, child in Q.grouped_prog.items():
        if isinstance(child, Plate):
            sample_plate(child, (*active, name), scope, noise, K, mode, dtype, result)
            continue
        if datagroup(child):
            continue
        K_axis = Kname(name)
        all_args = set(a for d in child.values() for a in d.all_args) - set(child.keys()) - {'prev'}
        for a in all_args:
            if a not in scope:
                raise Exception(f"{a} is not in scope")
        local = resample_scope({k: v for k, v in scope.items() if k in all_args}, name, K_axis, noise, mode)
        ts_perm = None
        if any(isinstance(d, Timeseries) for d in child.values()):
            ts_perm = perm_from_uniform(noise[(name, 'timeseries')], mode)
        for var, d in child.items():
            e = noise[var]
            eps = ONT(e.to(dtype), tuple(active) + (K_axis,))
            if isinstance(d, Timeseries):
                x = timeseries_draw(d, var, local, active, K_axis, eps, ts_perm, dtype)
            else:
                x = draw(d, local, eps, dtype)
            local[var] = x
            scope[var] = x
            result[var] = NT(x.t, x.axes)


def sample_q(Q: Plate, inputs_params: dict, noise: dict, K: int, mode: int = 0, dtype=t.float32) -> dict:
    """BoundPlate._sample (BoundPlate.py:338-363) with explicit noise.  Returns {varname: NT[plates..., K, *event]}."""
    scope = {k: ONT(v.t.to(dtype) if v.t.is_floating_point() else v.t, v.axes) for k, v in (inputs_params or {}).items()}
    result = {}
    sample_plate(Q, (), scope, noise, K, mode, dtype, result)
    return result
