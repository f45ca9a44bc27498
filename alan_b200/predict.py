"""Prediction: the step AFTER posterior resampling.

Mirror of the reference's `Sample.importance_sample(N)` -> `ImportanceSample.extend(...)` ->
`ExtendedImportanceSample.predictive_ll(data)` (reference src/alan/ImportanceSample.py:28-177,
Plate.py:145-215, dist.py:234-294):

  * `extend`: every variable of the prior P (latents AND data) is drawn over the EXTENDED plates with the posterior
    sample axis N as its sample axis -- ancestrally, each draw conditioned on the extended draws before it -- and the
    original block (the posterior sample of a latent, the observed values of a data variable) is pasted back into the
    leading corner of the extended tensor (dist.py:247-269);
  * `predictive_ll`: log p(extended data | extended sample) per data variable, summed over its plates for all cells and
    for the training block, `logmeanexp_N(all - train)` (ImportanceSample.py:152-177).

Both are single programs of the same engine the logPQ path runs on (factor-VM draws / densities, `PasteOp`, fixed-order
reductions, `LSE_eps`), executed with one C-ABI call each.  Randomness is explicit base noise, as in
alan_b200/sampling.py: parity with the reference is defined for identical noise (tests/test_gpu_predict.py).
Timeseries are not extended here (raises).
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from .model import Plate, Dist, Data, Timeseries
from .named import NT
from .plan import (Planner, TensorSig, PT, PasteOp, ReduceOp, LeafRef, plain, R_SUM, R_LSE_EPS, _prod)
from .sampling import NOISE_KIND, _draw_expr
from .trace import Expr, _bshape


def _walk_vars(P: Plate, active=()):
    """(varname, dist, active plates) in the reference's order (Plate.sample_extended walks flat_prog)."""
    for name, child in P.flat_prog.items():
        if isinstance(child, Plate):
            yield from _walk_vars(child, (*active, name))
        else:
            yield name, child, tuple(active)


class _Program:
    """Shared plumbing: a bare planner over extended plate sizes + the axis N, inputs in a fixed order."""

    def __init__(self, sizes, canon, sig, dtype):
        self.pl = Planner.bare(sig, sizes, dtype, canon)
        self.dtype = dtype

    def finish(self):
        plan = self.pl.plan
        plan.programs = [self.pl.fwd]
        plan.n_fwd, plan.n_bwd = 1, 0
        plan.assign_offsets(self.pl.itemsize)
        plan.serialize()
        self.plan = plan
        self.dp = None


class Extender(_Program):
    """The program of `ImportanceSample.extend` for one (P, shapes, N)."""

    def __init__(self, P: Plate, samples: dict, data: dict, ext_sizes: dict, ext_inputs: dict, N: int, dtype, device=None):
        all_plates = P.all_platenames()
        sizes = {a: int(ext_sizes[a]) for a in all_plates}
        sizes['N'] = int(N)
        canon = list(all_plates) + ['N']
        order = lambda axes: tuple(a for a in canon if a in axes)
        sig = {}
        self.in_axes = {}
        for k, v in (ext_inputs or {}).items():
            self.in_axes[k] = order(v.axes)
            sig[k] = TensorSig('param', self.in_axes[k], v.pos_shape)
            for a, n in v.named_sizes.items():
                if sizes.get(a) != n:
                    raise Exception(f"extended input {k} has {n} elements along {a}; the extended plate size is {sizes.get(a)}")
        super().__init__(sizes, canon, sig, dtype)
        pl = self.pl
        self.device = device
        # the ORIGINAL tensors keep their own (smaller) plate sizes: explicit inputs, read only by the paste ops
        self.orig = {}
        self.orig_order = []
        for k, v in {**data, **samples}.items():
            axes = order(v.axes)
            osz = dict(v.named_sizes)
            for a in axes:
                if a != 'N' and osz[a] > sizes[a]:
                    raise Exception(f"extended plate {a} ({sizes[a]}) is smaller than the original ({osz[a]})")
            name = f"__orig_{k}"
            pt = PT(axes, v.pos_shape, {**sizes, **osz}, 'input', index=len(pl.plan.input_names), name=name)
            pl.plan.input_names.append(name)
            pl.plan.input_pts[name] = pt
            self.orig[k] = (pt, axes)
            self.orig_order.append((name, k, axes))
        self.noise, self.outputs = [], []
        scope = {k: Expr.leaf(pl.inputs[k], s.axes, s.pos_shape) for k, s in sig.items()}
        for var, d, active in _walk_vars(P):
            if isinstance(d, Timeseries):
                raise Exception("extend: Timeseries are not supported")
            if isinstance(d, Data):
                raise Exception(f"{var}: the prior P cannot contain Data()")
            axes = tuple(active) + ('N',)
            args = {k: pl.resolve_arg(d.family, k, v, scope) for k, v in d.args.items()}
            shape = ()
            for a in args.values():
                shape = _bshape(shape, a.pos_shape)
            if d.family not in NOISE_KIND:
                _draw_expr(d.family, args, Expr.const(0.0))          # raises with the list of supported families
            nname = f"__noise_{var}"
            npt = pl._add_input(nname, axes, shape)
            pl.sig[nname] = TensorSig('param', axes, shape)
            self.noise.append((var, NOISE_KIND[d.family], axes, tuple(shape), nname))
            body = pl._prepare(_draw_expr(d.family, args, Expr.leaf(npt, axes, shape)))
            out = PT(axes, shape, pl.sizes, 'output', index=len(self.outputs), name=var)
            pl.emit_expr(body, nred=0, tag=f'extend:{var}', out=out)
            self.outputs.append((var, axes, tuple(shape)))
            if var in self.orig:
                # the original block goes back into the leading corner; observed data carry no N axis: broadcast
                src, saxes = self.orig[var]
                if not set(saxes) <= set(axes) or tuple(src.pos_shape) != tuple(shape):
                    raise Exception(f"{var}: the original tensor has axes {saxes} + {src.pos_shape}, the prior draws {axes} + {shape}")
                ss, ds = src.cstrides(), out.cstrides()
                dims = []
                for i, a in enumerate(axes):
                    if a in saxes:
                        j = saxes.index(a)
                        dims.append((src.shape[j], ss[j], ds[i]))
                    else:
                        dims.append((out.shape[i], 0, ds[i]))
                for e in range(len(shape)):
                    dims.append((shape[e], ss[len(saxes) + e], ds[len(axes) + e]))
                pl.fwd.append(PasteOp(src, out, dims))
            scope[var] = Expr.leaf(out, axes, shape)
        self.finish()

    def noise_shapes(self):
        return {var: (kind, tuple([self.pl.sizes[a] for a in axes] + list(pos))) for var, kind, axes, pos, _ in self.noise}

    def make_noise(self, device, seed=None):
        g = torch.Generator(device=device)
        g.manual_seed(int(seed)) if seed is not None else g.seed()
        return {var: (torch.randn if kind == 'normal' else torch.rand)(shape, dtype=self.dtype, device=device, generator=g)
                for var, (kind, shape) in self.noise_shapes().items()}

    def run(self, samples: dict, data: dict, ext_inputs: dict, noise=None, seed=None) -> dict:
        from . import runtime
        if self.dp is None:
            self.dp = runtime.DevicePlan(self.plan, self.device)
        dev = self.dp.device
        if noise is None:
            noise = self.make_noise(dev, seed)
        by_input = {}
        for var, kind, axes, pos, name in self.noise:
            x = noise[var]
            want = tuple([self.pl.sizes[a] for a in axes] + list(pos))
            if tuple(x.shape) != want:
                raise Exception(f"noise for {var}: expected shape {want}, got {tuple(x.shape)}")
            by_input[name] = x.to(dev).to(self.dtype).contiguous()
        src = {**data, **samples}
        for name, k, axes in self.orig_order:
            by_input[name] = src[k].order(axes).t.detach().to(dev).to(self.dtype).contiguous()
        ins = []
        for name in self.plan.input_names:
            if name in self.plan.const_inputs:
                ins.append(self.dp.consts[name])
            elif name in by_input:
                ins.append(by_input[name])
            else:
                ins.append(ext_inputs[name].order(self.in_axes[name]).t.detach().to(dev).to(self.dtype).contiguous())
        outs = [torch.empty([self.pl.sizes[a] for a in axes] + list(pos), dtype=self.dtype, device=dev)
                for _, axes, pos in self.outputs]
        self.dp.run(0, ins, outs)
        return {var: NT(o, axes) for (var, axes, _), o in zip(self.outputs, outs)}


class PredictiveLL(_Program):
    """The program of `ExtendedImportanceSample.predictive_ll` for one (P, shapes, N): one output scalar per data
    variable."""

    def __init__(self, P: Plate, ext_samples: dict, ext_data: dict, orig_sizes: dict, ext_inputs: dict, N: int, dtype, device=None):
        all_plates = P.all_platenames()
        sizes = {}
        for v in list(ext_samples.values()) + list(ext_data.values()) + list((ext_inputs or {}).values()):
            sizes.update({a: n for a, n in v.named_sizes.items()})
        sizes['N'] = int(N)
        canon = list(all_plates) + ['N']
        order = lambda axes: tuple(a for a in canon if a in axes)
        sig, self.in_axes = {}, {}
        for kind, d in (('param', ext_inputs or {}), ('sample', ext_samples), ('data', ext_data)):
            for k, v in d.items():
                if k in sig:
                    continue                                         # a data variable also present among the samples
                self.in_axes[k] = order(v.axes)
                sig[k] = TensorSig(kind if kind != 'data' else 'param', self.in_axes[k], v.pos_shape)
        super().__init__(sizes, canon, sig, dtype)
        pl = self.pl
        self.device = device
        scope = {k: Expr.leaf(pl.inputs[k], s.axes, s.pos_shape) for k, s in sig.items() if k not in ext_data}
        self.vars = []
        for var, d, active in _walk_vars(P):
            if var not in ext_data:
                continue
            value = Expr.leaf(pl.inputs[var], sig[var].axes, sig[var].pos_shape)
            F = pl.density(d, value, scope, tag=f'll:{var}')           # [plates..., N]
            plates = [a for a in F.axes if a != 'N']
            nd = [pl.axdim('N')]
            tot = pl.ws(('N',), name=f'll_all:{var}')
            pl.emit(ReduceOp(R_SUM, tot, nd, [pl.axdim(a) for a in plates], [(plain(F), 1.0)], tag=f'sum_all:{var}'))
            trn = pl.ws(('N',), name=f'll_train:{var}')
            rd_train = [('ax', a, int(orig_sizes[a])) for a in plates]
            pl.emit(ReduceOp(R_SUM, trn, nd, rd_train, [(plain(F), 1.0)], tag=f'sum_train:{var}'))
            out = PT((), (), pl.sizes, 'output', index=len(self.vars), name=f'pll:{var}')
            pl.emit(ReduceOp(R_LSE_EPS, out, [], nd, [(plain(tot), 1.0), (plain(trn), -1.0)], cadd=-math.log(N), tag=f'logmeanexp_N:{var}'))
            self.vars.append(var)
        self.finish()

    def run(self, ext_samples: dict, ext_data: dict, ext_inputs: dict) -> dict:
        from . import runtime
        if self.dp is None:
            self.dp = runtime.DevicePlan(self.plan, self.device)
        dev = self.dp.device
        src = {**(ext_inputs or {}), **ext_samples, **ext_data}
        ins = []
        for name in self.plan.input_names:
            if name in self.plan.const_inputs:
                ins.append(self.dp.consts[name])
            else:
                ins.append(src[name].order(self.in_axes[name]).t.detach().to(dev).to(self.dtype).contiguous())
        outs = [torch.empty((), dtype=self.dtype, device=dev) for _ in self.vars]
        self.dp.run(0, ins, outs)
        return dict(zip(self.vars, outs))


class AbstractImportanceSample(dict):
    """N joint posterior samples: {varname: NT with axes ('N', plates...)} (the dict IS `dump()`)."""

    def dump(self) -> dict:
        return dict(self)

    def moments(self, specs):
        """[(varname | tuple of varnames, f)] -> [NT]: the mean over N of f(samples) (moments.py:13-14 from_samples)."""
        out = []
        for v, f in specs:
            vs = (v,) if isinstance(v, str) else tuple(v)
            xs = [self[x] for x in vs]
            plates = []
            for x in xs:
                plates += [a for a in x.axes if a != 'N' and a not in plates]
            N = xs[0].named_sizes['N']
            ts = []
            for x in xs:
                y = x.order(('N',) + tuple(a for a in plates if a in x.axes)).t
                shape = [N] + [x.named_sizes.get(a, 1) for a in plates] + list(x.pos_shape)
                ts.append(y.reshape(shape))
            out.append(NT(f(*ts).mean(0), tuple(plates)))
        return out


class ImportanceSample(AbstractImportanceSample):
    """Returned by `Sample.importance_sample(N)` (reference ImportanceSample.py:28-98)."""

    def __init__(self, problem, samples: dict, N: int):
        super().__init__(samples)
        self.problem, self.N = problem, int(N)

    def extend(self, extended_platesizes: dict, extended_inputs: Optional[dict] = None, noise=None, seed=None):
        from .problem import _as_nt
        p = self.problem
        if not isinstance(extended_platesizes, dict):
            raise Exception("extended_platesizes must be a dict {plate name: size}")
        ext_sizes = dict(extended_platesizes)
        for a, n in p.platesizes.items():
            ext_sizes.setdefault(a, n)
        if set(ext_sizes) != set(p.platesizes):
            raise Exception(f"extended_platesizes names plates {sorted(set(ext_sizes) - set(p.platesizes))} the model does not have")
        ext_inputs = {k: _as_nt(v) for k, v in (extended_inputs or {}).items()}
        missing = set(p.inputs) - set(ext_inputs)
        if missing:
            raise Exception(f"the model has inputs {sorted(missing)}: their extended versions must be given to extend()")
        dtype = torch.float64 if any(v.t.dtype == torch.float64 for v in self.values()) else torch.float32
        key = ('extend', tuple(sorted(ext_sizes.items())), self.N, dtype,
               tuple(sorted((k, v.axes, tuple(v.t.shape)) for k, v in {**self, **ext_inputs}.items())))
        if key not in p._runners:
            p._runners[key] = Extender(p.P, dict(self), p.data, ext_sizes, ext_inputs, self.N, dtype, p.device)
        ext = p._runners[key].run(dict(self), p.data, ext_inputs, noise=noise, seed=seed)
        out = {k: NT(v.order(('N',) + tuple(a for a in v.axes if a != 'N')).t.contiguous(), ('N',) + tuple(a for a in v.axes if a != 'N'))
               for k, v in ext.items()}
        return ExtendedImportanceSample(p, out, self.N, ext_sizes, ext_inputs)


class ExtendedImportanceSample(AbstractImportanceSample):
    """Returned by `ImportanceSample.extend` (reference ImportanceSample.py:100-177)."""

    def __init__(self, problem, samples: dict, N: int, ext_sizes: dict, ext_inputs: dict):
        super().__init__(samples)
        self.problem, self.N, self.ext_sizes, self.ext_inputs = problem, int(N), ext_sizes, ext_inputs

    def predictive_ll(self, data: dict) -> dict:
        """{data variable: 0-d tensor} average predictive log-likelihood of the test cells; `data` holds ALL the data
        (train + test) of every variable that was extended, the others are taken from the problem."""
        from .problem import _as_nt
        p = self.problem
        if not isinstance(data, dict):
            raise Exception("data must be a dict {variable name: tensor}")
        ext_data = {k: _as_nt(v) for k, v in data.items()}
        extra = set(ext_data) - set(p.data)
        if extra:
            raise Exception(f"{sorted(extra)} are not data variables of the problem")
        for k, v in ext_data.items():
            for a, n in v.named_sizes.items():
                if n != self.ext_sizes.get(a):
                    raise Exception(f"extended data {k} has {n} elements along {a}; the extended plate size is {self.ext_sizes.get(a)}")
        dtype = torch.float64 if any(v.t.dtype == torch.float64 for v in self.values()) else torch.float32
        latents = {k: v for k, v in self.items() if k not in p.data}
        key = ('pll', tuple(sorted(self.ext_sizes.items())), self.N, dtype,
               tuple(sorted((k, v.axes, tuple(v.t.shape)) for k, v in {**latents, **ext_data, **self.ext_inputs}.items())))
        if key not in p._runners:
            p._runners[key] = PredictiveLL(p.P, latents, ext_data, p.platesizes, self.ext_inputs, self.N, dtype, p.device)
        return p._runners[key].run(latents, ext_data, self.ext_inputs)
