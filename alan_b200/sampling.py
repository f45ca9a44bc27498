"""Ancestral sampling of Q on the GPU: the step BEFORE the logPQ path.

Mirror of the reference's `Problem.sample(K)` -> `BoundPlate._sample` -> `Plate.sample` -> `sample_gdt` ->
`Sampler.resample_scope` / `Dist.sample` / `Timeseries.sample` (reference src/alan/Problem.py:71-97,
BoundPlate.py:338-363, Plate.py:93-143, dist.py:23-72, Sampler.py:85-169, Timeseries.py:89-123):

  * every latent group draws K particles per plate cell; the particles of its parents (each parent has its own K axis)
    are first PERMUTED along K independently per plate cell of the parent (PermutationSampler: argsort of uniforms) or
    resampled (CategoricalSampler: uniform picks), so that child particle k is conditioned on parent particle
    perm[k] -- the mixture proposal `reduce_logQ` later averages over (Sampler.py:118-134);
  * a Timeseries is drawn step by step, each step conditioned on the previous step's particles permuted by
    `timeseries_perm[t]`.

The walk happens once per (model, shapes, K): it emits ONE program -- permutation, gather and factor-VM expression ops
(csrc/sampling.cuh, csrc/kernels.cuh) -- executed with a single C-ABI call.  Randomness enters as EXPLICIT base noise:
float64 uniforms for the permutations, standard normals / uniforms in the working dtype for the draws
(`loc + scale * eps`, inverse CDFs), generated on the device by torch by default or supplied by the caller -- which is
how sampling parity is defined (tests/test_sampling_*.py: identical samples for identical base noise; the reference's
own RNG stream depends on torchdim's internal dim order).

Families with a closed-form transform are drawn natively: Normal, LogNormal, HalfNormal, Exponential, Uniform,
Laplace, Bernoulli, MultivariateNormal (loc + scale_tril eps, the factor from csrc/mvn.cuh).  The rejection-sampled ones (Gamma, Beta, StudentT, Poisson, Binomial, NegativeBinomial) raise.
"""
from __future__ import annotations

import math

import torch

from .model import Plate, Dist, Data, Timeseries, datagroup, Kname
from .named import NT
from .plan import (Planner, TensorSig, PT, PermOp, KGatherOp, TsSampleOp, LeafRef, plain, _prod)
from .trace import Expr, _bshape


class PermutationSampler:
    """Permute the parent particles: each parent particle has exactly one child (reference Sampler.py:139-148)."""
    mode = 0


class CategoricalSampler:
    """Resample the parent particles with a uniform categorical (reference Sampler.py:150-160)."""
    mode = 1


class IndependentSampler:
    """K independent draws from the full joint: child particle k is conditioned on parent particle k (reference
    Sampler.py:162-169; the sampler of `Problem.sample_nonmp`).  No permutation tensor: the parent is read under the
    child's K axis name."""
    mode = 2


NOISE_KIND = {'Normal': 'normal', 'LogNormal': 'normal', 'HalfNormal': 'normal', 'Exponential': 'uniform',
              'Uniform': 'uniform', 'Laplace': 'uniform', 'Bernoulli': 'uniform', 'MultivariateNormal': 'normal'}


def _draw_expr(family, args, noise: Expr) -> Expr:
    """The draw as an expression of the distribution arguments and one base-noise tensor."""
    mk = Expr.make
    if family == 'Normal':
        return mk('add', args['loc'], mk('mul', args['scale'], noise))
    if family == 'LogNormal':
        return mk('exp', mk('add', args['loc'], mk('mul', args['scale'], noise)))
    if family == 'HalfNormal':
        return mk('mul', mk('abs', noise), args['scale'])
    if family == 'Exponential':
        return mk('div', mk('neg', mk('log1p', mk('neg', noise))), args['rate'])
    if family == 'Uniform':
        return mk('add', args['low'], mk('mul', mk('sub', args['high'], args['low']), noise))
    if family == 'Laplace':
        s = mk('sub', noise, Expr.const(0.5))
        return mk('sub', args['loc'], mk('mul', mk('mul', args['scale'], mk('div', s, mk('abs', s))),
                                          mk('log1p', mk('neg', mk('mul', Expr.const(2.0), mk('abs', s))))))
    if family == 'Bernoulli':
        p = args['probs'] if 'probs' in args else mk('sigmoid', args['logits'])
        return mk('lt', noise, p)
    raise Exception(f"sampling from {family} on the device is not supported (no closed-form transform of base noise); "
                    f"supported: {sorted(NOISE_KIND)}")


class QSampler:
    """Sampling program for one (Q, shapes, K).  `run(inputs_params, noise=None, seed=None)` -> {varname: NT}."""

    def __init__(self, Q: Plate, inputs_params: dict, platesizes: dict, K: int, sampler=PermutationSampler,
                 dtype=torch.float32, device=None):
        self.Q, self.K, self.sampler, self.dtype = Q, int(K), sampler, dtype
        self.platesizes = dict(platesizes)
        groups = Q.groupvarnames()
        all_plates = Q.all_platenames()
        for a in all_plates:
            if a not in self.platesizes:
                raise Exception(f"the size of plate {a} is unknown: pass platesizes={{'{a}': ...}}")
        sizes = dict(self.platesizes)
        for g in groups:
            sizes[Kname(g)] = self.K
        self.canon = list(all_plates) + [Kname(g) for g in groups]
        sig = {}
        self.param_order = []
        for k, v in (inputs_params or {}).items():
            axes = tuple(a for a in self.canon if a in v.axes)
            sig[k] = TensorSig('param', axes, v.pos_shape)
            self.param_order.append((k, axes))
        self.pl = Planner.bare(sig, sizes, dtype, self.canon)
        self.noise = []              # [(key, kind, axes, pos_shape, input name)] in plan input order
        self.outputs = []            # [(varname, axes, pos_shape)] = program outputs
        self.scope = {k: Expr.leaf(self.pl.inputs[k], s.axes, s.pos_shape) for k, s in sig.items()}
        self.var_pt = {}
        self._walk(Q, (), dict(self.scope))
        plan = self.pl.plan
        plan.programs = [self.pl.fwd]
        plan.n_fwd, plan.n_bwd = 1, 0
        plan.assign_offsets(self.pl.itemsize)
        plan.serialize()
        self.plan = plan
        self.dp = None
        self.device = device

    # ------------------------------------------------------------------ plan construction
    def _noise_input(self, key, kind, axes, pos_shape):
        name = f"__noise{len(self.noise)}"
        pt = self.pl._add_input(name, axes, pos_shape)
        self.pl.sig[name] = TensorSig('param', axes, pos_shape)
        self.noise.append((key, kind, tuple(axes), tuple(pos_shape), name))
        return pt

    def _perm(self, key, plates, K_axis):
        """perm[plates..., K] over the dims of a parent (its plates and its K axis), from float64 uniforms."""
        pl = self.pl
        rows = _prod(pl.sizes[a] for a in plates)
        # float64 uniforms whatever the working dtype: the input is typed by the caller, the plan only sees a pointer
        u = self._noise_input(key, 'perm', tuple(plates) + (K_axis,), ())
        out = pl.ws_raw(rows * self.K * (8 // pl.itemsize), name=f'perm:{key}')
        pl.fwd.append(PermOp(u, out, rows, self.K, self.sampler.mode))
        return out

    def _walk(self, Q: Plate, active, scope):
        pl = self.pl
        for name, child in Q.grouped_prog.items():
            if isinstance(child, Plate):
                self._walk(child, (*active, name), dict(scope))
                continue
            if datagroup(child):
                continue
            Kg = Kname(name)
            all_args = set(a for d in child.values() for a in d.all_args) - set(child.keys()) - {'prev'}
            for a in all_args:
                if a not in scope:
                    raise Exception(f"{a} is not in scope")
            # ---- resample_scope: permute every parent along ITS K axis, per cell of ITS plates (Sampler.py:85-116)
            local = {}
            by_K = {}
            for a in [k for k in scope if k in all_args]:                 # scope order = insertion order, as upstream
                e = scope[a]
                ks = [x for x in e.axes if x.startswith('K_')]
                if len(ks) > 1:
                    raise Exception(f"{a} carries several K axes")
                by_K.setdefault(ks[0] if ks else None, []).append(a)
            for Kp, names in by_K.items():
                if Kp is None:
                    for a in names:
                        local[a] = scope[a]
                    continue
                e0 = scope[names[0]]
                plates0 = tuple(x for x in e0.axes if x != Kp)
                if self.sampler.mode == 2:                                 # IndependentSampler: perm = arange
                    for a in names:
                        e = scope[a]
                        if e.op != 'leaf' or e.rename or e.mode:
                            raise Exception(f"internal: {a} is not a plain tensor")
                        local[a] = Expr.leaf(e.ref, tuple(Kg if x == Kp else x for x in e.axes), e.pos_shape,
                                             rename={Kg: Kp})
                    continue
                perm = self._perm((name, Kp), plates0, Kp)
                for a in names:
                    e = scope[a]
                    if tuple(x for x in e.axes if x != Kp) != plates0:
                        raise Exception(f"{a} and {names[0]} share {Kp} but not their plates")
                    outer = _prod(pl.sizes[x] for x in plates0)
                    inner = _prod(e.pos_shape)
                    out = pl.ws(plates0 + (Kg,), e.pos_shape, name=f'resampled:{a}')
                    pl.fwd.append(KGatherOp(e.ref, perm, out, outer, self.K, inner))
                    local[a] = Expr.leaf(out, out.axes, out.pos_shape)
            ts_perm = None
            if any(isinstance(d, Timeseries) for d in child.values()) and self.sampler.mode == 2:
                raise Exception("IndependentSampler does not draw a Timeseries here (SampleNonMP has no Timeseries support "
                                "upstream either, SampleNonMP.py:156)")
            if any(isinstance(d, Timeseries) for d in child.values()):
                ts_perm = self._perm((name, 'timeseries'), tuple(active), Kg)
            # ---- the draws, in program order; later members of a Group see the earlier ones (dist.py:60-70)
            for var, d in child.items():
                axes = tuple(active) + (Kg,)
                if isinstance(d, Timeseries):
                    out = self._timeseries(var, d, active, Kg, local, ts_perm)
                elif d.family in ('MultivariateNormal', 'LowRankMultivariateNormal'):
                    # loc + L eps with L = scale_tril from the device factorisation (torch's rsample)
                    loc, L, _, _, dd = pl.mvn_parts(d, local)
                    npt = self._noise_input(var, 'normal', axes, (dd,))
                    eps = Expr.leaf(npt, axes, (dd,))
                    body = pl._prepare(Expr.make('add', loc, Expr.make('sumlast', Expr.make('mul', L, eps))))
                    out = PT(axes, (dd,), pl.sizes, 'output', index=len(self.outputs), name=var)
                    pl.emit_expr(body, nred=0, tag=f'draw:{var}', out=out)
                    self.outputs.append((var, axes, (dd,)))
                else:
                    args = {k: pl.resolve_arg(d.family, k, v, local) for k, v in d.args.items()}
                    shape = ()
                    for a in args.values():
                        shape = _bshape(shape, a.pos_shape)
                    if d.family not in NOISE_KIND:
                        _draw_expr(d.family, args, Expr.const(0.0))
                    npt = self._noise_input(var, NOISE_KIND[d.family], axes, shape)
                    body = pl._prepare(_draw_expr(d.family, args, Expr.leaf(npt, axes, shape)))
                    out = PT(axes, shape, pl.sizes, 'output', index=len(self.outputs), name=var)
                    pl.emit_expr(body, nred=0, tag=f'draw:{var}', out=out)
                    self.outputs.append((var, axes, tuple(shape)))
                e = Expr.leaf(out, out.axes, out.pos_shape)
                local[var] = e
                scope[var] = e
                self.var_pt[var] = out

    def _timeseries(self, var, ts: Timeseries, active, Kg, local, ts_perm):
        """Timeseries.sample (Timeseries.py:89-123): x_t ~ trans(prev = permuted x_{t-1}), x_{-1} = the resampled init."""
        pl = self.pl
        if not active:
            raise Exception("a Timeseries must live in a plate (its time axis)")
        T_axis = active[-1]
        d = ts.trans
        if ts.init not in local:
            raise Exception(f"Timeseries initial state {ts.init} is not in scope")
        init = local[ts.init]
        if set(init.axes) != set(active[:-1]) | {Kg}:
            raise Exception(f"Initial state, {ts.init}, doesn't have the right dimensions for timeseries {var}; the "
                            f"initial state must be defined one step up in the plate heirarchy")
        axes = tuple(active) + (Kg,)
        prev_pt = pl.ws(axes, init.pos_shape, name=f'prev:{var}')          # never read from memory: the kernel feeds it
        sc = {**local, 'prev': Expr.leaf(prev_pt, axes, init.pos_shape)}
        args = {k: pl.resolve_arg(d.family, k, v, sc) for k, v in d.args.items()}
        shape = init.pos_shape
        for a in args.values():
            shape = _bshape(shape, a.pos_shape)
        if shape != init.pos_shape:
            raise Exception("the transition of a Timeseries must keep the shape of its initial state")
        npt = self._noise_input(var, NOISE_KIND.get(d.family) or _draw_expr(d.family, args, Expr.const(0.)), axes, shape)
        body = pl._prepare(_draw_expr(d.family, args, Expr.leaf(npt, axes, shape)))
        out = PT(axes, shape, pl.sizes, 'output', index=len(self.outputs), name=var)
        self.outputs.append((var, axes, tuple(shape)))
        _, eop = pl.emit_expr(body, nred=0, tag=f'draw:{var}', out=out, append=False)
        leaves = eop.codeobj.leaves
        prev_idx = [i for i, lf in enumerate(leaves) if lf.pt is prev_pt]
        if len(prev_idx) > 1:
            raise Exception("internal: `prev` read through several leaves")
        dims = eop.keep
        keys = [(x[0], x[1]) for x in dims]
        if keys[:len(axes)] != [('ax', a) for a in axes]:
            raise Exception("internal: Timeseries draw is not laid out [plates..., T, K, event]")
        # the initial state must be materialised as [outer plates..., K, event]: it is (a resampled tensor or a draw)
        init_pt = init.ref
        n_outer = _prod(pl.sizes[a] for a in active[:-1])
        E = _prod(shape)
        pl.fwd.append(TsSampleOp(eop, prev_idx[0] if prev_idx else -1, len(active) - 1, len(active), init_pt, ts_perm,
                                 n_outer, pl.sizes[T_axis], self.K, E))
        return out

    # ------------------------------------------------------------------ execution
    def noise_shapes(self):
        """{key: (kind, shape, dtype)} of the base noise a call consumes (keys: variable names for the draws,
        (group, parent K axis | 'timeseries') for the permutations)."""
        out = {}
        for key, kind, axes, pos, _ in self.noise:
            shape = [self.pl.sizes[a] for a in axes] + list(pos)
            out[key] = (kind, tuple(shape), torch.float64 if kind == 'perm' else self.dtype)
        return out

    def make_noise(self, device, seed=None):
        g = torch.Generator(device=device)
        if seed is not None:
            g.manual_seed(int(seed))
        else:
            g.seed()
        out = {}
        for key, (kind, shape, dt) in self.noise_shapes().items():
            if kind == 'normal':
                out[key] = torch.randn(shape, dtype=dt, device=device, generator=g)
            else:
                out[key] = torch.rand(shape, dtype=dt, device=device, generator=g)
        return out

    def run(self, inputs_params: dict, noise=None, seed=None) -> dict:
        from . import runtime
        if self.dp is None:
            self.dp = runtime.DevicePlan(self.plan, self.device)
        dev = self.dp.device
        if noise is None:
            noise = self.make_noise(dev, seed)
        shapes = self.noise_shapes()
        noise_by_input = {}
        for key, kind, axes, pos, name in self.noise:
            x = noise[key]
            kind_, shape, dt = shapes[key]
            if tuple(x.shape) != shape:
                raise Exception(f"noise for {key}: expected shape {shape}, got {tuple(x.shape)}")
            noise_by_input[name] = x.to(dev).to(dt).contiguous()
        ins = []
        for name in self.plan.input_names:
            if name in self.plan.const_inputs:
                ins.append(self.dp.consts[name])
            elif name in noise_by_input:
                ins.append(noise_by_input[name])
            else:
                axes = next(a for k, a in self.param_order if k == name)
                ins.append(inputs_params[name].order(axes).t.detach().to(dev).to(self.dtype).contiguous())
        outs = [torch.empty([self.pl.sizes[a] for a in axes] + list(pos), dtype=self.dtype, device=dev)
                for _, axes, pos in self.outputs]
        self.dp.run(0, ins, outs)
        return {var: NT(o, axes) for (var, axes, _), o in zip(self.outputs, outs)}
