"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the logPQ plate-tree reduction.

A plain-PyTorch (no functorch.dim, no CUDA) restatement of the reference's
algorithm for the hot path, function by function.  It is the checker for the
CUDA engine: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it.  The product (alan_b200/) never does.

Parity pin: tests/golden/*.pt hold outputs of the UNMODIFIED reference run in
the build container (tests/golden/make_golden.py); tests/test_oracle_golden.py
checks this file against every one of them (fp32 and fp64).

Reference map (all paths relative to /root/reference/):
  lse_eps / logmeanexp      src/alan/utils.py:207-225
  dist_log_prob             src/alan/dist.py:211-232,297-302; src/alan/TorchDimDist.py:127-162
  timeseries_log_prob       src/alan/Timeseries.py:203-245
  reduce_logQ               src/alan/Sampler.py:118-134
  logPQ_gdt                 src/alan/logpq.py:157-254
  lp_getter                 src/alan/logpq.py:257-332
  collect_lps / reduce_Ks   src/alan/reduce_Ks.py:236-298
  chain_logmmexp            src/alan/utils.py:478-510
  logPQ_plate               src/alan/logpq.py:15-155 (Split: src/alan/Split.py:44-130)
  sample_Ks / logPQ_sample  src/alan/reduce_Ks.py:35-83; src/alan/sample_logpq.py:17-107
  index_into_sample         src/alan/Sample.py:359-381
  marginals / moments       src/alan/Sample.py:208-272,291-346

Deviations from the reference, all deliberate and documented in DESIGN.md:
  * every number/constant is converted to the working dtype (the reference makes
    them float32 0-d tensors, dist.py:311-318, which leaks fp32 rounding into
    fp64 runs); fp64 goldens are generated under torch.set_default_dtype(float64).
  * the contraction order comes from alan_b200.path.greedy_path because the
    reference's opt_einsum is absent; the golden generator gives
    the reference the same rule through oracle/shims/opt_einsum.
  * resampling takes EXPLICIT float64 uniforms and applies the inverse-CDF rule
    in float64 instead of torch.multinomial's RNG stream.
"""
from __future__ import annotations

import math
import types
import numbers

import torch as t
import torch.distributions as td

from alan_b200.model import Plate, Dist, Data, Timeseries, datagroup, Kname, function_arguments
from alan_b200.named import NT
from alan_b200.path import greedy_path


# ---------------------------------------------------------------------------
# Named-axis tensor with torch-function support so that model lambdas such as
# ``lambda z, x: z @ x`` or ``lambda a: t.exp(a)`` run on it unchanged.
# ---------------------------------------------------------------------------

def _align(xs):
    """Align ONTs to common named axes (union, first-appearance order) with
    positional dims right-aligned.  Returns (raw tensors/scalars, axes)."""
    axes = []
    for x in xs:
        if isinstance(x, ONT):
            for a in x.axes:
                if a not in axes:
                    axes.append(a)
    P = max([x.t.ndim - len(x.axes) for x in xs if isinstance(x, ONT)], default=0)
    out = []
    for x in xs:
        if not isinstance(x, ONT):
            out.append(x)
            continue
        raw = x.t
        p = raw.ndim - len(x.axes)
        perm = [x.axes.index(a) for a in axes if a in x.axes] + list(range(len(x.axes), raw.ndim))
        raw = raw.permute(perm)
        shape = []
        it = iter(raw.shape[:len(x.axes)])
        for a in axes:
            shape.append(next(it) if a in x.axes else 1)
        shape += [1] * (P - p) + list(raw.shape[len(x.axes):])
        out.append(raw.reshape(shape))
    return out, tuple(axes)


class ONT:
    def __init__(self, tensor, axes):
        self.t = tensor
        self.axes = tuple(axes)
        assert tensor.ndim >= len(self.axes)

    # -- structure ---------------------------------------------------------
    @property
    def pos_ndim(self):
        return self.t.ndim - len(self.axes)

    def sizes(self):
        return {a: int(s) for a, s in zip(self.axes, self.t.shape)}

    def order(self, axes):
        """named dims `axes` first (in that order), remaining named dims after."""
        axes = tuple(axes)
        rest = tuple(a for a in self.axes if a not in axes)
        new = axes + rest
        perm = [self.axes.index(a) for a in new] + list(range(len(self.axes), self.t.ndim))
        return ONT(self.t.permute(perm), new)

    def sum_pos(self):
        """sum_non_dim: sum every positional dim (reference utils.py:147-152)."""
        if self.pos_ndim == 0:
            return self
        return ONT(self.t.sum(tuple(range(len(self.axes), self.t.ndim))), self.axes)

    def reduce(self, fn, axes):
        axes = tuple(a for a in axes)
        if not axes:
            return self
        dims = tuple(self.axes.index(a) for a in axes)
        return ONT(fn(self.t, dims), tuple(a for a in self.axes if a not in axes))

    def sum(self, axes):
        return self.reduce(lambda x, d: x.sum(d), axes)

    def amax(self, axes):
        return self.reduce(lambda x, d: x.amax(d), axes)

    def index_axis(self, axis, idx: "ONT"):
        """x.order(axis)[idx]: replace named `axis` by the named axes of the integer
        tensor idx (reference sample_logpq.py:75-77, reduce_Ks.py:55-56).  Axes shared
        between x and idx are matched elementwise (advanced-indexing semantics of
        first-class dims)."""
        assert idx.pos_ndim == 0
        x = self.order((axis,))
        rest = x.axes[1:]
        out_axes = tuple(idx.axes) + tuple(a for a in rest if a not in idx.axes)
        # build index tensors for every dim of x.t, broadcast over out_axes (+ positional)
        sizes = {**self.sizes(), **idx.sizes()}
        npos = x.pos_ndim
        full = [sizes[a] for a in out_axes]

        def view_for(a_list_axes, tensor):
            shape = [sizes[a] if a in a_list_axes else 1 for a in out_axes] + [1] * npos
            perm = [a_list_axes.index(a) for a in out_axes if a in a_list_axes]
            return tensor.permute(perm).reshape(shape)
        index = [view_for(list(idx.axes), idx.t)]
        for a in rest:
            ar = t.arange(sizes[a])
            index.append(view_for([a], ar))
        for k in range(npos):
            shape = [1] * (len(out_axes) + npos)
            shape[len(out_axes) + k] = x.t.shape[len(x.axes) + k]
            index.append(t.arange(x.t.shape[len(x.axes) + k]).reshape(shape))
        res = x.t[tuple(index)]
        return ONT(res, out_axes)

    # -- arithmetic --------------------------------------------------------
    @classmethod
    def __torch_function__(cls, func, types_, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in (t.matmul, t.Tensor.matmul, t.Tensor.__matmul__):
            return _matmul(args[0], args[1])
        if func is t.Tensor.__rmatmul__:
            return _matmul(args[1], args[0])
        flat = list(args)
        raw, axes = _align(flat)
        res = func(*raw, **kwargs)
        return ONT(res, axes)

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        fn = getattr(t.Tensor, name)

        def method(*args, **kwargs):
            return ONT.__torch_function__(fn, (ONT,), (self, *args), kwargs)
        return method

    def _bin(self, other, fn):
        return ONT.__torch_function__(fn, (ONT,), (self, other))

    def __add__(self, o): return self._bin(o, t.add)
    def __radd__(self, o): return ONT.__torch_function__(t.add, (ONT,), (o, self))
    def __sub__(self, o): return self._bin(o, t.sub)
    def __rsub__(self, o): return ONT.__torch_function__(lambda a, b: a - b, (ONT,), (o, self))
    def __mul__(self, o): return self._bin(o, t.mul)
    def __rmul__(self, o): return ONT.__torch_function__(t.mul, (ONT,), (o, self))
    def __truediv__(self, o): return self._bin(o, t.div)
    def __rtruediv__(self, o): return ONT.__torch_function__(lambda a, b: a / b, (ONT,), (o, self))
    def __pow__(self, o): return self._bin(o, t.pow)
    def __neg__(self): return ONT(-self.t, self.axes)
    def __matmul__(self, o): return _matmul(self, o)
    def __rmatmul__(self, o): return _matmul(o, self)


def _matmul(a, b):
    """``@`` on the positional dims with named dims as batch (how functorch.dim treats it)."""
    if not isinstance(a, ONT):
        a = ONT(t.as_tensor(a), ())
    if not isinstance(b, ONT):
        b = ONT(t.as_tensor(b), ())
    pa, pb = a.pos_ndim, b.pos_ndim
    if pa == 1 and pb == 1:
        (ra, rb), axes = _align([a, b])
        return ONT((ra * rb).sum(-1), axes)
    if pa == 2 and pb == 1:
        b2 = ONT(b.t.unsqueeze(-1), b.axes)
        (ra, rb), axes = _align([a, b2])
        return ONT(t.matmul(ra, rb).squeeze(-1), axes)
    if pa == 1 and pb == 2:
        a2 = ONT(a.t.unsqueeze(-2), a.axes)
        (ra, rb), axes = _align([a2, b])
        return ONT(t.matmul(ra, rb).squeeze(-2), axes)
    if pa == 2 and pb == 2:
        (ra, rb), axes = _align([a, b])
        return ONT(t.matmul(ra, rb), axes)
    raise Exception("oracle: unsupported matmul ranks")


def ont(x: NT) -> ONT:
    return ONT(x.t, x.axes)


# ---------------------------------------------------------------------------
# A1  LSE with eps  (utils.py:207-225)
# ---------------------------------------------------------------------------

# eps of A1/A5 is finfo(dtype).eps (utils.py:220,507).  `EPS_OF` lets a test evaluate the fp32 SEMANTICS (fp32's eps)
# in float64 arithmetic: the yardstick for fp32 rounding error (tests/golden_io.py f64_truth).
EPS_OF = None


def _eps(dtype):
    return t.finfo(EPS_OF if EPS_OF is not None else dtype).eps


def lse_eps(x: ONT, axes) -> ONT:
    axes = tuple(a for a in axes if a in x.axes)          # ignore_extra_dims=True (reduce_Ks.py:251)
    if len(axes) == 0:
        return x
    x_max = x.amax(axes)
    s = (x - x_max).exp().sum(axes)
    return (s + _eps(s.t.dtype)).log() + x_max


def logmeanexp(x: ONT, axes) -> ONT:
    sizes = x.sizes()
    return lse_eps(x, axes) - sum([math.log(sizes[a]) for a in axes])


# ---------------------------------------------------------------------------
# A2  density  (dist.py:211-232, TorchDimDist.py:127-162)
# ---------------------------------------------------------------------------

def resolve_arg(v, scope, dtype):
    """dist.py:211-229 (paramname2val): number / tensor / scope string / lambda."""
    if isinstance(v, str):
        return scope[v]
    if isinstance(v, types.FunctionType):
        val = v(*[scope[a] for a in function_arguments(v)])
        if not isinstance(val, ONT):
            raise Exception("Lambda on a distribution returned a non-Tensor")
        return val
    if isinstance(v, t.Tensor):
        return ONT(v.to(dtype), ())
    assert isinstance(v, numbers.Number)
    return ONT(t.tensor(float(v), dtype=dtype), ())


def dist_log_prob(dist: Dist, value: ONT, scope: dict, dtype) -> ONT:
    """log_prob on fully broadcast operands, then sum of EVERY positional dim
    (TorchDimDist.py:157-162)."""
    args = {k: resolve_arg(v, scope, dtype) for k, v in dist.args.items()}
    names = list(args.keys())
    if dist.family == 'MultivariateNormal':
        # event-aware alignment (TorchDimDist.py:44-62): loc / value are vectors, the matrix argument has event_dim 2
        mname = [k for k in names if k != 'loc'][0]
        (rv, rloc, rmat_as_vec), axes = _align([value, args['loc'], ONT(args[mname].t[..., 0], args[mname].axes)])
        mat = args[mname]
        perm = [mat.axes.index(a) for a in axes if a in mat.axes] + [len(mat.axes), len(mat.axes) + 1]
        shape = [mat.sizes()[a] if a in mat.axes else 1 for a in axes] + list(mat.t.shape[-2:])
        rmat = mat.t.permute(perm).reshape(shape)
        d = td.MultivariateNormal(rloc, validate_args=False, **{mname: rmat})
        return ONT(d.log_prob(rv), axes)
    if dist.family in ('Categorical', 'RelaxedOneHotCategorical', 'LowRankMultivariateNormal'):
        # event-aware alignment (TorchDimDist.py:44-77): every argument is laid out [unnamed batch dims, named axes,
        # its own event dims] with event_dim taken from the torch distribution's constraints, the value likewise with
        # the support's event_dim; what is left after log_prob are the named axes and the unnamed batch dims
        D = getattr(td, dist.family)
        # (torch lists no constraint for `temperature`, a scalar per cell -- which is why the reference itself cannot
        # construct the Relaxed* families: KeyError at TorchDimDist.py:47)
        ev = [D.support.event_dim] + [D.arg_constraints[k].event_dim if k in D.arg_constraints else 0 for k in names]
        xs = [value] + [args[k] for k in names]
        axes = []
        for x in xs:
            for a in x.axes:
                if a not in axes:
                    axes.append(a)
        nb = [x.t.ndim - len(x.axes) - e for x, e in zip(xs, ev)]
        B = max(nb)
        raw = []
        for x, e, b in zip(xs, ev, nb):
            perm = [x.axes.index(a) for a in axes if a in x.axes] + list(range(len(x.axes), x.t.ndim))
            r = x.t.permute(perm)
            pos = list(r.shape[len(x.axes):])
            it = iter(r.shape[:len(x.axes)])
            shape = [next(it) if a in x.axes else 1 for a in axes] + [1] * (B - b) + pos
            raw.append(r.reshape(shape))
        d = D(**dict(zip(names, raw[1:])), validate_args=False)
        return ONT(d.log_prob(raw[0]), tuple(axes)).sum_pos()
    raw, axes = _align([value] + [args[k] for k in names])
    rv, rargs = raw[0], dict(zip(names, raw[1:]))
    d = getattr(td, dist.family)(**rargs, validate_args=False)
    lp = d.log_prob(rv)
    return ONT(lp, axes).sum_pos()


def timeseries_log_prob(ts: Timeseries, sample: ONT, scope: dict, T_axis, K_axis, dtype):
    """Timeseries.py:203-245: prev = concat(init, x[:-1]) along T with K -> Kinit."""
    init = scope[ts.init]
    diff = [a for a in init.axes if a not in sample.axes]
    assert len(diff) == 1
    Kinit = diff[0]
    x = sample.order((T_axis, K_axis))                     # [T, K, rest..., pos...]
    rest = x.axes[2:]
    prev_body = ONT(x.t[:-1], (T_axis, Kinit) + rest)      # K renamed to Kinit
    init_o = init.order((Kinit,) + tuple(a for a in rest if a in init.axes))
    # broadcast init over any axes of `rest` it lacks, then prepend it along T
    tgt_axes = (Kinit,) + rest
    sizes = {**sample.sizes(), **init.sizes()}
    init_full = _expand_to(init, tgt_axes, sizes, x.t.shape[2 + len(rest):])
    prev = t.cat([init_full.unsqueeze(0), prev_body.t], 0)
    prev = ONT(prev, (T_axis,) + tgt_axes)
    scope = {**scope, "prev": prev}
    lp = dist_log_prob(ts.trans, sample, scope, dtype)
    return lp, Kinit


def _expand_to(x: ONT, axes, sizes, pos_shape):
    xo = x.order(tuple(a for a in axes if a in x.axes))
    shape = [sizes[a] if a in xo.axes else 1 for a in axes]
    p = xo.pos_ndim
    raw = xo.t.reshape(shape + [1] * (len(pos_shape) - p) + list(xo.t.shape[len(xo.axes):]))
    return raw.expand([sizes[a] for a in axes] + list(pos_shape)).contiguous()


# ---------------------------------------------------------------------------
# A3  group factor  (logpq.py:157-254, Sampler.py:118-134)
# ---------------------------------------------------------------------------

def reduce_logQ(lq: ONT, active_plates, K_axis) -> ONT:
    parents = tuple(a for a in lq.axes if a != K_axis and a not in active_plates)
    return logmeanexp(lq, parents)


def logPQ_gdt(name, prog_P, prog_Q, sample, data, scope, active_plates, varname2groupvarname, dtype):
    if datagroup(prog_Q):
        assert len(prog_Q) == 1
        k = next(iter(prog_Q))
        lp = dist_log_prob(prog_P[k], data[k], scope, dtype)
        return lp, (), (), ()

    K_axis = Kname(name)
    K = None
    total_logP, total_logQ = 0., 0.
    T_axis = active_plates[-1] if active_plates else None
    Kinits = []
    for k in prog_P:
        dP, dQ, x = prog_P[k], prog_Q[k], sample[k]
        K = x.sizes()[K_axis]
        lp, kin = (timeseries_log_prob(dP, x, scope, T_axis, K_axis, dtype) if isinstance(dP, Timeseries)
                   else (dist_log_prob(dP, x, scope, dtype), None))
        lq, _ = (timeseries_log_prob(dQ, x, scope, T_axis, K_axis, dtype) if isinstance(dQ, Timeseries)
                 else (dist_log_prob(dQ, x, scope, dtype), None))
        if kin is not None:
            Kinits.append(kin)
        total_logP = lp + total_logP
        total_logQ = lq + total_logQ
    total_logQ = reduce_logQ(total_logQ, active_plates, K_axis)
    lp = total_logP - total_logQ - math.log(K)
    if Kinits:
        return lp, (), (K_axis,), (Kinits[0],)
    return lp, (K_axis,), (), ()


# ---------------------------------------------------------------------------
# A4  contraction  (reduce_Ks.py:236-298)
# ---------------------------------------------------------------------------

def collect_lps(lps, Ks_to_sum):
    sizes = {}
    for lp in lps:
        sizes.update(lp.sizes())
    path = greedy_path([lp.axes for lp in lps], Ks_to_sum, sizes)
    all_reduced = [list(lps)]
    Ks_to_sample = []
    lps = list(lps)
    for idxs in path:
        chosen = [lps[i] for i in idxs]
        lps = [lps[i] for i in range(len(lps)) if i not in idxs]
        remaining_axes = set(a for lp in lps for a in lp.axes)
        chosen_axes = []
        for lp in chosen:
            for a in lp.axes:
                if a not in chosen_axes:
                    chosen_axes.append(a)
        # deterministic order (the reference iterates a Python set here, reduce_Ks.py:276)
        ks = tuple(k for k in Ks_to_sum if k in chosen_axes and k not in remaining_axes)
        Ks_to_sample.append(ks)
        s = chosen[0]
        for c in chosen[1:]:
            s = s + c
        lps.append(lse_eps(s, ks))
        all_reduced.append(list(lps))
    all_reduced = all_reduced[:-1]
    assert len(lps) == 1
    keep = [i for i, ks in enumerate(Ks_to_sample) if ks != ()]
    return lps[0], [all_reduced[i] for i in keep], [Ks_to_sample[i] for i in keep], path


def reduce_Ks(lps, Ks_to_sum):
    return collect_lps(lps, Ks_to_sum)[0]


# ---------------------------------------------------------------------------
# A5  chain  (utils.py:478-510)
# ---------------------------------------------------------------------------

def logmmexp(prev, curr):
    prev_max = prev.amax(-1, keepdim=True)
    curr_max = curr.amax(-2, keepdim=True)
    r = (prev - prev_max).exp() @ (curr - curr_max).exp()
    return (r + _eps(r.dtype)).log() + prev_max + curr_max


def chain_logmmexp(ms):
    """ms: [T, ..., K, K] with T leading (batch dims in the middle)."""
    while ms.shape[0] != 1:
        prev, curr = ms[::2], ms[1::2]
        rem = None
        if len(prev) > len(curr):
            rem = prev[-1:]
            prev = prev[:-1]
        ms = logmmexp(prev, curr)
        if rem is not None:
            ms = t.cat([ms, rem], 0)
    return ms[0]


# ---------------------------------------------------------------------------
# a1/a2  plate recursion  (logpq.py:15-155, 257-332)
# ---------------------------------------------------------------------------

class Ctx:
    def __init__(self, sample, inputs_params, data, extra_log_factors, all_plates, dtype, split=None,
                 checkpoint=False):
        self.sample = sample
        self.inputs_params = inputs_params
        self.data = data
        self.elf = extra_log_factors
        self.all_plates = tuple(all_plates)
        self.dtype = dtype
        self.split = split
        self.checkpoint = checkpoint      # the reference's `checkpoint` strategy (logpq.py:41,62-66)


def _elf_at_level(elf, active_plates, all_plates):
    """extra_log_factors live at the plate whose active plates equal the plate axes they
    carry (tensordict2tree, Plate.py:355-377)."""
    out = []
    for k, v in elf.items():
        plates = set(a for a in v.axes if a in all_plates)
        if plates == set(active_plates):
            out.append(v)
    return out


def lp_getter(name, P, Q, ctx, scope, active_plates, v2g):
    lps = _elf_at_level(ctx.elf, active_plates, ctx.all_plates)
    Knon, Kts, Kinits = [], [], []
    for childname, childQ in Q.grouped_prog.items():
        if isinstance(childQ, dict):
            childP = {v: P.flat_prog[v] for v in childQ}
            lp, a, b, c = logPQ_gdt(childname, childP, childQ, ctx.sample, ctx.data, scope, active_plates,
                                    v2g, ctx.dtype)
        else:
            lp = logPQ_plate(childname, P.flat_prog[childname], childQ, ctx, scope, active_plates, v2g)
            a = b = c = ()
        lps.append(lp)
        Knon.extend(a); Kts.extend(b); Kinits.extend(c)
    return lps, Knon, Kts, Kinits


def _slice_ctx(ctx, plate, lo, hi):
    def sl(d):
        out = {}
        for k, v in d.items():
            if plate in v.axes:
                i = v.axes.index(plate)
                out[k] = ONT(v.t.narrow(i, lo, hi - lo), v.axes)
            else:
                out[k] = v
        return out
    return Ctx(sl(ctx.sample), sl(ctx.inputs_params), sl(ctx.data), sl(ctx.elf), ctx.all_plates, ctx.dtype,
               ctx.split, ctx.checkpoint)


def split_sizes(orig, size):
    """Split.py:84-95"""
    sizes = [size] * (orig // size)
    if orig % size:
        sizes.append(orig % size)
    if size > 2 and sizes[-1] == 1:
        sizes[-2] -= 1
        sizes[-1] += 1
    return sizes


def logPQ_plate(name, P, Q, ctx, scope, active_plates, v2g):
    if ctx.split is not None and ctx.split[0] == name:
        plate, size = ctx.split
        n = None
        for d in (ctx.sample, ctx.data, ctx.inputs_params, ctx.elf):
            for v in d.values():
                if plate in v.axes:
                    n = v.sizes()[plate]
        lpq, lo = None, 0
        for s in split_sizes(n, size):
            sub = _slice_ctx(ctx, plate, lo, lo + s)
            sub_scope = {k: (ONT(v.t.narrow(v.axes.index(plate), lo, s), v.axes) if plate in v.axes else v)
                         for k, v in scope.items()}
            if ctx.checkpoint and t.is_grad_enabled():
                # logpq.py:62-66: every Split chunk under a non-reentrant torch.utils.checkpoint, so the backward
                # holds one chunk's intermediates at a time (cfg-5: 200 chunks of 50 users)
                from torch.utils.checkpoint import checkpoint as _ckpt
                axes_box = []

                def run(prev_t, sub=sub, sub_scope=sub_scope, lpq=lpq):
                    prev = None if lpq is None else ONT(prev_t, lpq.axes)
                    r = _logPQ_plate(name, P, Q, sub, sub_scope, active_plates, v2g, prev)
                    axes_box[:] = [r.axes]
                    return r.t
                out = _ckpt(run, None if lpq is None else lpq.t, use_reentrant=False)
                lpq = ONT(out, axes_box[0])
            else:
                lpq = _logPQ_plate(name, P, Q, sub, sub_scope, active_plates, v2g, lpq)
            lo += s
        return lpq
    return _logPQ_plate(name, P, Q, ctx, scope, active_plates, v2g, None)


def _logPQ_plate(name, P, Q, ctx, scope, active_plates, v2g, prev_lpq):
    if name is not None:
        active_plates = [*active_plates, name]
    scope = {**scope}
    for d in (ctx.inputs_params, ctx.sample):
        for k, v in d.items():
            scope[k] = v
    lps, all_Ks, K_currs, K_inits = lp_getter(name, P, Q, ctx, scope, active_plates, v2g)
    lp = reduce_Ks(lps, all_Ks)
    if name is not None:
        if len(K_inits) > 0:
            o = lp.order((name, *K_inits, *K_currs))             # [T, Kinit, Kcurr, rest...]
            rest = o.axes[3:]
            nb = len(rest)
            # move batch axes between T and the two K axes: [T, rest..., Kp, Kc]
            perm = [0] + list(range(3, 3 + nb)) + [1, 2]
            ms = o.t.permute(perm)
            r = chain_logmmexp(ms)                               # [rest..., Kp, Kc]
            r = t.logsumexp(r, -1)                               # no eps (logpq.py:139)
            lp = ONT(r, rest + (K_inits[0],))
            assert prev_lpq is None
        else:
            lp = lp.sum((name,))
            if prev_lpq is not None:
                lp = prev_lpq + lp
    return lp


def elbo(P: Plate, Q: Plate, sample, inputs_params, data, extra_log_factors=None, split=None, checkpoint=False):
    """Sample._elbo (Sample.py:69-108).  All dict values are NT/ONT; returns a 0-d tensor.
    split=(plate, size) is `computation_strategy=Split(plate, size)` (Split.py:44-130); checkpoint=True wraps
    every chunk in torch.utils.checkpoint like the reference's default strategy (logpq.py:41,62-66)."""
    conv = lambda d: {k: (v if isinstance(v, ONT) else ont(v)) for k, v in (d or {}).items()}
    sample, inputs_params, data, elf = conv(sample), conv(inputs_params), conv(data), conv(extra_log_factors)
    elf = {k: v.sum_pos() for k, v in elf.items()}                # Sample.py:74
    dtype = _working_dtype(sample, inputs_params, data, elf)
    sample, inputs_params, data, elf = [_cast(d, dtype) for d in (sample, inputs_params, data, elf)]
    ctx = Ctx(sample, inputs_params, data, elf, P.all_platenames(), dtype, split, checkpoint)
    lp = logPQ_plate(None, P, Q, ctx, {}, [], Q.varname2groupvarname())
    assert lp.t.ndim == 0, f"elbo has leftover axes {lp.axes}"
    return lp.t


def _working_dtype(*dicts):
    dt = t.float32
    for d in dicts:
        for v in d.values():
            if v.t.dtype == t.float64:
                dt = t.float64
    return dt


def _cast(d, dtype):
    return {k: (ONT(v.t.to(dtype), v.axes) if v.t.is_floating_point() else v) for k, v in d.items()}


# ---------------------------------------------------------------------------
# A7  marginals and moments via source terms  (Sample.py:208-272, 291-346)
# ---------------------------------------------------------------------------

def marginals(P, Q, sample, inputs_params, data, joints=(), split=None):
    """Returns {frozenset(groupvarnames): NT with axes (K..., plates...)}."""
    g2p = Q.groupvarname2platenames()
    sizes = {}
    for d in (sample, data, inputs_params):
        for v in d.values():
            sizes.update(v.named_sizes if isinstance(v, NT) else v.sizes())
    keys = [frozenset([g]) for g in Q.groupvarnames()] + [frozenset(j) for j in joints]
    dtype = _working_dtype({k: ont(v) if isinstance(v, NT) else v for k, v in sample.items()})
    Js, elf = [], {}
    for key in keys:
        gs = tuple(sorted(key, key=lambda g: Q.groupvarnames().index(g)))
        axes = tuple(Kname(g) for g in gs) + tuple(g2p[gs[0]])
        J = t.zeros([sizes[a] for a in axes], dtype=dtype, requires_grad=True)
        Js.append(J)
        elf[key] = ONT(J, axes)
    sample_d = {k: v.detach() for k, v in sample.items()}
    L = elbo(P, Q, sample_d, inputs_params, data, elf, split)
    grads = t.autograd.grad(L, Js)
    return {key: NT(g, elf[key].axes) for key, g in zip(keys, grads)}


def moments(P, Q, sample, inputs_params, data, moms, split=None):
    """moms: list of (varnames tuple, f).  Returns list of NT [plates..., *f.shape]
    (Sample.py:291-346)."""
    all_plates = P.all_platenames()
    dtype = _working_dtype({k: ont(v) for k, v in sample.items()})
    Js, elf, axes_l = [], {}, []
    for i, (varnames, f) in enumerate(moms):
        xs = [ont(sample[v].detach()) for v in varnames]
        fx = f(*xs)
        if not isinstance(fx, ONT):
            raise Exception("moment function must return a tensor")
        plates = tuple(a for a in all_plates if a in fx.axes)
        fx = fx.order(plates)
        psz = fx.sizes()
        J = t.zeros([psz[a] for a in plates] + list(fx.t.shape[len(fx.axes):]), dtype=dtype, requires_grad=True)
        Js.append(J)
        axes_l.append(plates)
        elf[i] = ONT(fx.t.detach().to(dtype), fx.axes) * ONT(J, plates)
    sample_d = {k: v.detach() for k, v in sample.items()}
    L = elbo(P, Q, sample_d, inputs_params, data, elf, split)
    grads = t.autograd.grad(L, Js)
    return [NT(g, a) for g, a in zip(grads, axes_l)]


# ---------------------------------------------------------------------------
# A8  posterior resampling  (sample_logpq.py:17-107, reduce_Ks.py:35-83)
# ---------------------------------------------------------------------------

def inverse_cdf_draw(lp: ONT, kaxes, u: ONT, N_axis="N") -> dict:
    """One joint categorical draw per (batch cell, n).

    lp: log-factor with named axes = batch axes (plates, possibly N) + kaxes.
    u : float64 uniforms with axes = batch axes of lp (without N) + (N,).
    Rule: p_j = exp(lp_j - max_j) in the factor dtype (reduce_Ks.py:62-66), then float64: j row-major over kaxes;
    c = cumsum(p);
    index = first j with c_j >= u * c_last (clamped to the last category).
    Returns {kaxis: integer ONT with axes (N, batch...)}.
    """
    batch = tuple(a for a in lp.axes if a not in kaxes)
    if N_axis not in batch:
        # broadcast over N
        lp = ONT(lp.t.unsqueeze(0).expand(u.sizes()[N_axis], *lp.t.shape), (N_axis,) + lp.axes)
        batch = (N_axis,) + batch
    b_noN = tuple(a for a in batch if a != N_axis)
    order = (N_axis,) + b_noN
    x = lp.order(order + tuple(kaxes))
    ksz = [x.sizes()[k] for k in kaxes]
    raw = x.t.reshape(*[x.sizes()[a] for a in order], -1)
    uu = u.order(order).t.to(t.float64)
    m = raw.amax(-1, keepdim=True)
    p = (raw - m).exp().to(t.float64)
    c = p.cumsum(-1)
    thr = (uu * c[..., -1]).unsqueeze(-1)
    flat = (c < thr).sum(-1).clamp(max=raw.shape[-1] - 1)
    out = {}
    rem = flat
    for k, sz in zip(reversed(kaxes), reversed(ksz)):       # unravel_index.py:97-100
        out[k] = ONT(rem % sz, order)
        rem = rem // sz
    return out


def sample_Ks(lps, Ks_to_sum, uniforms, N_axis, indices):
    """reduce_Ks.py:35-83 with explicit uniforms; `uniforms` is an iterator yielding, for each
    sampling step (in the order the steps are visited), a function batch_axes -> ONT."""
    _, lps_for_sampling, Ks_to_sample, _ = collect_lps(lps, Ks_to_sum)
    indices = dict(indices)
    new = {}
    for step_lps, kdims in zip(lps_for_sampling[::-1], Ks_to_sample[::-1]):
        lp = step_lps[0]
        for c in step_lps[1:]:
            lp = lp + c
        for a in list(lp.axes):
            if a in new:
                lp = lp.index_axis(a, new[a])
        batch = tuple(a for a in lp.axes if a not in kdims and a != N_axis)
        u = uniforms(kdims, batch)
        drawn = inverse_cdf_draw(lp, kdims, u, N_axis)
        new.update(drawn)
    return new


def logPQ_sample(name, P, Q, ctx, scope, active_plates, v2g, indices, uniforms, N_axis):
    if name is not None:
        active_plates = [*active_plates, name]
    scope = {**scope}
    for d in (ctx.inputs_params, ctx.sample):
        for k, v in d.items():
            scope[k] = v
    lps, non_ts, ts_Ks, ts_inits = lp_getter(name, P, Q, ctx, scope, active_plates, v2g)
    if len(ts_Ks) > 0:
        raise Exception("importance_sample through a Timeseries is unfinished in the reference "
                        "(README.md:41-44; reduce_Ks.py:223 fails)")
    lps = list(lps)
    for i in range(len(lps)):
        for a in list(lps[i].axes):
            if a in indices:
                lps[i] = lps[i].index_axis(a, indices[a])
    if len(non_ts) > 0:
        indices = {**indices, **sample_Ks(lps, non_ts, uniforms, N_axis, {})}
    for childname, childQ in Q.grouped_prog.items():
        if isinstance(childQ, Plate):
            indices = logPQ_sample(childname, P.flat_prog[childname], childQ, ctx, scope, active_plates, v2g,
                                   indices, uniforms, N_axis)
    return indices


def importance_sample_idxs(P, Q, sample, inputs_params, data, uniforms, N_axis="N"):
    """Sample._importance_sample_idxs (Sample.py:150-183).  `uniforms(kdims, batch_axes)` returns
    an ONT of float64 uniforms with axes batch_axes + (N,).  Returns {groupvarname: NT[N, plates...]}."""
    conv = lambda d: {k: (v if isinstance(v, ONT) else ont(v)) for k, v in (d or {}).items()}
    sample, inputs_params, data = conv(sample), conv(inputs_params), conv(data)
    dtype = _working_dtype(sample, inputs_params, data)
    sample, inputs_params, data = [_cast(d, dtype) for d in (sample, inputs_params, data)]
    ctx = Ctx(sample, inputs_params, data, {}, P.all_platenames(), dtype, None)
    with t.no_grad():
        idx = logPQ_sample(None, P, Q, ctx, {}, [], Q.varname2groupvarname(), {}, uniforms, N_axis)
    out = {}
    for g in Q.groupvarnames():
        out[g] = NT(idx[Kname(g)].t, idx[Kname(g)].axes)
    return out


def index_into_sample(sample, indices, Q: Plate):
    """Sample.py:359-381: value.order(Kdim)[indices[group]]."""
    v2g = Q.varname2groupvarname()
    out = {}
    for name, value in sample.items():
        g = v2g[name]
        idx = indices[g]
        r = ont(value.detach()).index_axis(Kname(g), ONT(idx.t, idx.axes))
        out[name] = NT(r.t, r.axes)
    return out
