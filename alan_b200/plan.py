"""Plan compiler: model tree + tensor signature -> static program for libalan_b200.so.

The reference walks the P/Q plate tree in Python on every call (src/alan/logpq.py:15-155,
257-332), builds torch.distributions objects, and lets autograd + checkpointing derive the
backward pass.  Here the walk happens ONCE per (model, shapes): it emits
  * a forward program  (factor expressions, log-semiring contractions, plate sums, chains),
  * a backward program (hand-derived adjoints of exactly those ops, pruned to what is needed),
  * a resampling program (top-down categorical draws over the retained factors),
as a flat list of ops over contiguous HBM tensors, serialised to an int32 blob that the C ABI
(include/alan_b200.h) executes with one host call per program.

Data layout in HBM: every tensor is contiguous row-major `[named axes..., positional dims...]`;
named axes are plates (program order) followed by K axes (program order) unless a kernel asks
for something else.  Factor tensors are materialised at their natural rank (cells), never at
cells x event (the reference's `[M,Kz,d,Kmu,Kpsi]` broadcast never exists).
"""
from __future__ import annotations

import math
import os
import struct
import types
import numbers
from dataclasses import dataclass, field
from typing import Optional

import torch

from .model import Plate, Dist, Data, Timeseries, datagroup, Kname, function_arguments, DISCRETE_ARGS
from .path import greedy_path
from .trace import Expr, Proxy, _as_proxy, trace_function, UNARY, BINARY

MAXD, MAXL, MAXI, MAXC, NREG = 10, 10, 32, 16, 32
MAGIC, VERSION = 0x0A1AB200, 1
SP_WS, SP_INPUT, SP_OUTPUT, SP_AUX = 0, 1, 2, 3
OP_FILL, OP_EXPR, OP_EXPR_BWD, OP_REDUCE, OP_CHAIN, OP_CHAIN_BWD, OP_SAMPLE, OP_NORMAL_FAN, OP_COPY, OP_DOT, \
    OP_FAN_LSE, OP_BERN_DOT, OP_FAN_BWD, OP_XREDUCE, OP_NORMAL_Q_BWD, OP_PERM, OP_KGATHER, OP_TS_SAMPLE, OP_DEPS, OP_NORMAL_POLY_SUM, OP_PASTE, \
    OP_MVN_PREP, OP_RSEQ = range(1, 24)
R_SUM, R_LSE_EPS, R_LSE, R_WSUM = 0, 1, 2, 3
HOIST_RATIO = 16
NONMP_K = 'K_'          # the one K axis every latent shares in a SampleNonMP plan (no group can be named '')

VOPS = {'load': 0, 'const': 1, 'add': 2, 'sub': 3, 'mul': 4, 'div': 5, 'neg': 6, 'exp': 7, 'log': 8, 'sigmoid': 9,
        'square': 10, 'sqrt': 11, 'reciprocal': 12, 'softplus': 13, 'tanh': 14, 'abs': 15, 'log1p': 16, 'pow': 17,
        'lgamma': 18, 'mov': 19, 'lt': 20, 'cos': 21, 'sin': 22,
        'Normal': 32, 'Bernoulli_logits': 33, 'Bernoulli_probs': 34, 'LogNormal': 35, 'Laplace': 36,
        'Exponential': 37, 'Gamma': 38, 'Beta': 39, 'Poisson': 40, 'Cauchy': 41, 'HalfNormal': 42, 'Uniform': 43,
        'StudentT': 44, 'NegativeBinomial_logits': 45, 'NegativeBinomial_probs': 46, 'Binomial_logits': 47,
        'Binomial_probs': 48}

# density op -> order of the distribution arguments after the value operand
DENSITY_ARGS = {
    'Normal': ('loc', 'scale'), 'LogNormal': ('loc', 'scale'), 'Laplace': ('loc', 'scale'),
    'Cauchy': ('loc', 'scale'), 'Exponential': ('rate',), 'Gamma': ('concentration', 'rate'),
    'Beta': ('concentration1', 'concentration0'), 'Poisson': ('rate',), 'HalfNormal': ('scale',),
    'Uniform': ('low', 'high'), 'StudentT': ('df', 'loc', 'scale'),
}


# Families whose log-density is COMPOSED from the VM's primitive operations (no opcode of their own): each entry is
# (argument order, function of Proxy operands restating torch.distributions.<family>.log_prob -- the file it follows is
# named on each -- with the transform chains of TransformedDistribution families written out).  Values outside the
# support are not masked to -inf (the reference runs with validate_args off and would return the same finite formula).
def _softplus(x):
    return Proxy(Expr.make('softplus', x.expr))


def _probs_to_logits(p):                         # torch distributions/utils.py probs_to_logits(is_binary=True)
    return p.log() - (-p).log1p()


def _lp_gumbel(x, loc, scale):                   # torch gumbel.py:68-71
    y = (loc - x) / scale
    return (y - y.exp()) - scale.log()


def _lp_weibull(x, scale, k):                    # weibull.py: Exponential(1) -> PowerTransform(1/k) -> AffineTransform(0, scale)
    x1 = x / scale
    x0 = x1.pow(k)
    return -x0 - ((x1 / x0) / k).abs().log() - scale.abs().log()


def _lp_pareto(x, scale, alpha):                 # pareto.py: Exponential(alpha) -> ExpTransform -> AffineTransform(0, scale)
    x0 = (x / scale).log()
    return (alpha.log() - alpha * x0) - x0 - scale.abs().log()


def _lp_halfcauchy(x, scale):                    # half_cauchy.py:68-75 over cauchy.py:78-84
    return -math.log(math.pi) - scale.log() - ((x / scale) ** 2).log1p() + math.log(2)


def _lp_chi2(x, df):                             # chi2.py: Gamma(0.5 df, 0.5); gamma.py:85-93
    c = df * 0.5
    return c * math.log(0.5) + (c - 1.0) * x.log() - x * 0.5 - c.lgamma()


def _lp_geometric(x, p):                         # geometric.py:114-121
    return x * (-p).log1p() + p.log()


def _lp_kumaraswamy(x, a, b):                    # kumaraswamy.py: Uniform(0,1) -> Power(1/b) -> Affine(1,-1) -> Power(1/a)
    return a.log() + b.log() + (a - 1.0) * x.log() + (b - 1.0) * (-(x.pow(a))).log1p()


def _lp_fisher_snedecor(x, df1, df2):            # fishersnedecor.py:88-98
    ct1, ct2, ct3 = df1 * 0.5, df2 * 0.5, df1 / df2
    t1 = (ct1 + ct2).lgamma() - ct1.lgamma() - ct2.lgamma()
    t2 = ct1 * ct3.log() + (ct1 - 1.0) * x.log()
    t3 = (ct1 + ct2) * (ct3 * x).log1p()
    return t1 + t2 - t3


def _lp_logit_relaxed_bernoulli(x, temperature, logits):       # relaxed_bernoulli.py:107-112
    diff = logits - x * temperature
    return temperature.log() + diff - 2.0 * diff.exp().log1p()


def _lp_relaxed_bernoulli(y, temperature, logits):             # relaxed_bernoulli.py:115-148: ... -> SigmoidTransform
    x = y.log() - (-y).log1p()
    return _lp_logit_relaxed_bernoulli(x, temperature, logits) + _softplus(-x) + _softplus(x)


def _lp_one_hot_categorical(v, logits):          # one_hot_categorical.py:104-108 over categorical.py:66-67,137-143
    return (v * logits).sum(-1) - logits.exp().sum(-1).log() * v.sum(-1)


def _lp_categorical(v, logits, iota):             # categorical.py:137-143: the class index selects one normalised logit
    lt = lambda a, b: Proxy(Expr.make('lt', a.expr, b.expr))
    onehot = 1.0 - lt(iota, v) - lt(v, iota)      # [j == v] for exact integers, over the last positional dim
    return (onehot * logits).sum(-1) - logits.exp().sum(-1).log()


def _lp_relaxed_one_hot_categorical(y, temperature, logits):    # relaxed_categorical.py:83-96 (ExpRelaxedCategorical) -> ExpTransform
    J = float(logits.expr.pos_shape[-1])
    x = y.log()
    score = logits - x * temperature
    lse = score.exp().sum(-1).log()
    return score.sum(-1) - J * lse + math.lgamma(J) + (J - 1.0) * temperature.log() - x.sum(-1)


def _lp_continuous_bernoulli(x, logits, p, hoist):              # continuous_bernoulli.py:160-205, lims = (0.499, 0.501)
    lt = lambda a, b: Proxy(Expr.make('lt', _as_proxy(a).expr, _as_proxy(b).expr))
    # torch.where(c, a, b) as c a + (1 - c) b with c in {0, 1}: both branches are finite by torch's own construction
    outside = hoist((1.0 - lt(0.499, p)) + lt(0.501, p))         # le(p, 0.499) | gt(p, 0.501)
    cut = hoist(outside * p + (1.0 - outside) * 0.499)
    below = 1.0 - lt(0.5, cut)                                   # le(cut, 0.5)
    cut_below = below * cut
    ge = 1.0 - lt(cut, 0.5)
    cut_above = ge * cut + (1.0 - ge)
    first = hoist(((-cut).log1p() - cut.log()).abs().log())
    second = hoist(below * (-2.0 * cut_below).log1p() + (1.0 - below) * (2.0 * cut_above - 1.0).log())
    log_norm = first - second
    xx = (p - 0.5) ** 2
    taylor = math.log(2.0) + (4.0 / 3.0 + 104.0 / 45.0 * xx) * xx
    norm = hoist(outside * log_norm + (1.0 - outside) * taylor)
    return Proxy(Expr.make('Bernoulli_logits', x.expr, logits.expr)) + norm


_I0_SMALL = [1.0, 3.5156229, 3.0899424, 1.2067492, 0.2659732, 0.360768e-1, 0.45813e-2]            # von_mises.py:14-22
_I0_LARGE = [0.39894228, 0.1328592e-1, 0.225319e-2, -0.157565e-2, 0.916281e-2, -0.2057706e-1, 0.2635537e-1,
             -0.1647633e-1, 0.392377e-2]


def _horner(y, coef):                            # von_mises.py _eval_poly
    res = coef[-1]
    for c in reversed(coef[:-1]):
        res = c + y * res
    return res


def _lp_von_mises(x, loc, kappa, hoist):          # von_mises.py:146-154 over _log_modified_bessel_fn (:47-71), order 0
    lt = lambda a, b: Proxy(Expr.make('lt', _as_proxy(a).expr, _as_proxy(b).expr))
    kappa = hoist(kappa)                          # the concentration itself (an expression of parameters / samples) once
    ys = kappa / 3.75
    small = hoist(_horner(ys * ys, _I0_SMALL).log())
    poly_large = hoist(_horner(3.75 / kappa, _I0_LARGE))
    large = hoist(kappa - 0.5 * kappa.log() + poly_large.log())
    is_small = lt(kappa, 3.75)
    log_i0 = hoist(is_small * small + (1.0 - is_small) * large)
    return kappa * (x - loc).cos() - math.log(2.0 * math.pi) - log_i0


def _lp_multinomial(v, logits):                  # multinomial.py:121-132 (total_count = 1: see model.Dist)
    norm = logits - logits.exp().sum(-1).log()
    return (v.sum(-1) + 1.0).lgamma() - (v + 1.0).lgamma().sum(-1) + (norm * v).sum(-1)


COMPOSED = {
    'Gumbel': (('loc', 'scale'), _lp_gumbel), 'Weibull': (('scale', 'concentration'), _lp_weibull),
    'Pareto': (('scale', 'alpha'), _lp_pareto), 'HalfCauchy': (('scale',), _lp_halfcauchy), 'Chi2': (('df',), _lp_chi2),
    'Geometric': (('probs',), _lp_geometric), 'Kumaraswamy': (('concentration1', 'concentration0'), _lp_kumaraswamy),
    'FisherSnedecor': (('df1', 'df2'), _lp_fisher_snedecor),
    'RelaxedBernoulli': (('temperature', 'logits'), _lp_relaxed_bernoulli),
    'OneHotCategorical': (('logits',), _lp_one_hot_categorical), 'Multinomial': (('logits',), _lp_multinomial),
    'Categorical': (('logits',), _lp_categorical),
    'RelaxedOneHotCategorical': (('temperature', 'logits'), _lp_relaxed_one_hot_categorical),
    'ContinuousBernoulli': (('logits', 'probs'), _lp_continuous_bernoulli),
    'VonMises': (('loc', 'concentration'), _lp_von_mises),
}
# how the missing one of (probs, logits) is obtained from the given one, per family
_VECTOR_FAMILIES = ('OneHotCategorical', 'Multinomial', 'Categorical', 'RelaxedOneHotCategorical')


# ----------------------------------------------------------------------------------------
# physical tensors and dims
# ----------------------------------------------------------------------------------------
class PT:
    """A contiguous tensor in HBM: `[axes..., pos...]` row-major."""
    _next = 0

    def __init__(self, axes, pos_shape, sizes, space, index=0, offset=0, name=''):
        self.id = PT._next
        PT._next += 1
        self.axes = tuple(axes)
        self.pos_shape = tuple(int(s) for s in pos_shape)
        self.shape = tuple(int(sizes[a]) for a in self.axes) + self.pos_shape
        self.space, self.index, self.offset, self.name = space, index, offset, name

    @property
    def numel(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    def cstrides(self):
        st, acc = [], 1
        for s in reversed(self.shape):
            st.append(acc)
            acc *= s
        return list(reversed(st))

    def __repr__(self):
        return f"PT#{self.id}<{self.name}:{self.axes}+{self.pos_shape}@{self.space}>"


@dataclass(frozen=True)
class LeafRef:
    """How an op reads a tensor: optional axis renaming (Timeseries prev) and shift mode."""
    pt: PT
    rename: tuple = ()          # ((name seen by op, axis name of pt), ...)
    mode: int = 0
    mdim: Optional[str] = None  # named axis the mode refers to

    def stride(self, dim):
        kind, key, size = dim
        st = self.pt.cstrides()
        if kind == 'ax':
            rn = dict(self.rename)
            if key not in rn and key in rn.values():
                return 0                       # the tensor's own axis is seen under another name here
            name = rn.get(key, key)
            if name in self.pt.axes:
                i = self.pt.axes.index(name)
                return st[i] if self.pt.shape[i] > 1 else 0
            return 0
        k = key                                # event dim, counted from the right
        n = len(self.pt.pos_shape)
        if k < n:
            i = len(self.pt.axes) + n - 1 - k
            return st[i] if self.pt.shape[i] > 1 else 0
        return 0


def plain(pt: PT) -> LeafRef:
    return LeafRef(pt)


# ----------------------------------------------------------------------------------------
# blob writer
# ----------------------------------------------------------------------------------------
class W:
    def __init__(self):
        self.w = []
        self.seen = {}          # id -> PT of every workspace tensor referenced while writing
        self.refs = []          # every tensor referenced, in order (dependency analysis: Plan.insert_deps)

    def i32(self, v):
        v = int(v)
        self.w.append(v if v < 2 ** 31 else v - 2 ** 32)

    def i64(self, v):
        v = int(v) & 0xFFFFFFFFFFFFFFFF
        self.i32(v & 0xFFFFFFFF)
        self.i32(v >> 32)

    def f64(self, v):
        self.i64(struct.unpack('<q', struct.pack('<d', float(v)))[0])

    def tref(self, pt: PT):
        self.refs.append(pt)
        if pt.space == 'ws':
            self.seen[pt.id] = pt
            self.i32(SP_WS); self.i64(pt.offset if pt.offset is not None else 0)
        elif pt.space == 'input':
            self.i32(SP_INPUT); self.i64(pt.index)
        elif pt.space == 'output':
            self.i32(SP_OUTPUT); self.i64(pt.index)
        elif pt.space == 'aux':
            self.i32(SP_AUX); self.i64(pt.index)
        else:
            raise Exception(f"bad space {pt.space}")


class Op:
    code = 0

    def payload(self, w: W):
        raise NotImplementedError

    def serialize(self, w: W):
        start = len(w.w)
        w.i32(self.code)
        w.i32(0)
        self.payload(w)
        w.w[start + 1] = len(w.w) - start


def _dims(w, a, b):
    if len(a) + len(b) > MAXD:
        raise Exception(f"op needs {len(a) + len(b)} iteration dims; the kernels support {MAXD}")
    w.i32(len(a)); w.i32(len(b))
    for d in list(a) + list(b):
        w.i32(d[2])


def _opnd(w, leaf: LeafRef, dims, with_mode):
    w.tref(leaf.pt)
    if with_mode:
        w.i32(leaf.mode)
        mdim = 0
        if leaf.mode:
            keys = [(d[0], d[1]) for d in dims]
            mdim = keys.index(('ax', leaf.mdim))
        w.i32(mdim)
    for d in dims:
        w.i64(leaf.stride(d))


def _coalesce(sizes, n_a, strides_list, pinned=()):
    """Drop extent-1 dims and merge adjacent dims that every operand walks contiguously
    (stride[i] == size[i+1] * stride[i+1]); never across the n_a boundary nor through `pinned`
    dims (shift / first-step dims).  Returns (sizes, n_a, strides_list, index map old->new)."""
    nd = len(sizes)
    keep = [i for i in range(nd) if sizes[i] > 1 or i in pinned]
    groups = []                       # list of lists of old dim indices
    for i in keep:
        if groups:
            j = groups[-1][-1]
            same_side = (j < n_a) == (i < n_a)
            ok = same_side and i not in pinned and j not in pinned and \
                all(st[j] == sizes[i] * st[i] for st in strides_list)
            if ok:
                groups[-1].append(i)
                continue
        groups.append([i])
    new_sizes, new_strides, remap = [], [[] for _ in strides_list], {}
    new_na = 0
    for g in groups:
        sz = 1
        for i in g:
            sz *= sizes[i]
            remap[i] = len(new_sizes)
        new_sizes.append(sz)
        if g[0] < n_a:
            new_na += 1
        for k, st in enumerate(strides_list):
            new_strides[k].append(st[g[-1]])
    return new_sizes, new_na, new_strides, remap


def _write_dims(w, sizes, n_a):
    if len(sizes) > MAXD:
        raise Exception(f"op needs {len(sizes)} iteration dims; the kernels support {MAXD}")
    w.i32(n_a); w.i32(len(sizes) - n_a)
    for sz in sizes:
        w.i32(sz)


def _strides_like(pt: PT, own_dims, dims):
    """strides of a contiguous tensor laid out over own_dims, seen from the op dims `dims`."""
    st, acc = {}, 1
    for d in reversed(own_dims):
        st[(d[0], d[1])] = acc if d[2] > 1 else 0
        acc *= d[2]
    return [st.get((d[0], d[1]), 0) for d in dims]


@dataclass
class Code:
    instrs: list
    consts: list
    res: int
    leaves: list            # LeafRef per leaf index

    def write(self, w: W):
        if len(self.instrs) > MAXI or len(self.consts) > MAXC:
            raise Exception("traced expression is too long for the factor VM "
                            f"({len(self.instrs)} instructions, {len(self.consts)} constants)")
        w.i32(len(self.instrs))
        for (op, dst, a, b, c, d) in self.instrs:
            w.i32(op | (dst << 8) | (a << 16) | (b << 24))
            w.i32(c | (d << 8))
        w.i32(len(self.consts))
        for c in self.consts:
            w.f64(c)
        w.i32(self.res)


class DepsOp(Op):
    """Not an operation: the dependency table of the program it opens (op index -> indices of the earlier ops it must
    follow).  The executor uses it while a program is captured into a CUDA graph to put independent ops on parallel
    branches (csrc/alan_b200.cu run_ops)."""
    code = OP_DEPS

    def __init__(self, deps):
        self.deps = deps

    def payload(self, w):
        w.i32(len(self.deps))
        for d in self.deps:
            w.i32(len(d))
            for j in d:
                w.i32(j)


class FillOp(Op):
    code = OP_FILL

    def __init__(self, pt, nbytes):
        self.pt, self.nbytes = pt, nbytes

    def payload(self, w):
        w.tref(self.pt); w.i64(self.nbytes)


class FillRegionOp(FillOp):
    """Zero the whole adjoint region of the workspace (resolved when offsets are assigned)."""
    def __init__(self, plan):
        self.plan = plan

    def payload(self, w):
        lo, hi = self.plan.adj_region
        w.i32(SP_WS); w.i64(lo); w.i64(max(hi - lo, 0))

    @property
    def pt(self):
        lo, hi = self.plan.adj_region
        return PT((), (1,), {}, 'ws', offset=lo)

    @property
    def nbytes(self):
        lo, hi = self.plan.adj_region
        return hi - lo


class XReduceOp(Op):
    """In-program cross-rank sum of small tensors over NVLink peer memory (csrc/kernels.cuh xreduce_body): the
    pieces are packed into this rank's symmetric buffer, the ranks exchange one flag each, every rank adds all
    ranks' packs IN RANK ORDER (bit-identical results everywhere) and unpacks in place.  Replaces the host-issued
    ncclAllReduce between program segments: the message is a few KB, i.e. pure latency."""
    code = OP_XREDUCE
    autodiff_as = 'skip'          # adjoint of (sum over ranks) w.r.t. the local partial = the replicated upstream adjoint

    def __init__(self, site, pieces):
        self.site, self.pieces = site, list(pieces)
        if len(self.pieces) > 16:
            raise Exception("more than 16 tensors in one cross-rank reduction")
        self.out = self.pieces[0]

    def payload(self, w):
        w.i32(self.site); w.i32(len(self.pieces))
        for pt in self.pieces:
            w.tref(pt); w.i64(pt.numel)


class ExprOp(Op):
    """out[keep] (+)= scale * sum_red VM(leaves)   (csrc/kernels.cuh expr_fwd_kernel)"""
    code = OP_EXPR

    def __init__(self, out, keep, red, code: Code, acc=0, scale=1.0, tag=''):
        self.out, self.keep, self.red, self.codeobj, self.acc, self.scale, self.tag = out, keep, red, code, acc, scale, tag

    def payload(self, w):
        w.tref(self.out); w.i32(self.acc); w.f64(self.scale)
        dims = self.keep + self.red
        leaves = self.codeobj.leaves
        if len(leaves) > MAXL:
            raise Exception("factor expression reads more than 10 tensors")
        keys = [(d[0], d[1]) for d in dims]
        pinned = {keys.index(('ax', lf.mdim)) for lf in leaves if lf.mode}
        strides = [[lf.stride(d) for d in dims] for lf in leaves]
        sizes, n_a, strides, remap = _coalesce([d[2] for d in dims], len(self.keep), strides, pinned)
        _write_dims(w, sizes, n_a)
        w.i32(len(leaves))
        for lf, st in zip(leaves, strides):
            w.tref(lf.pt); w.i32(lf.mode)
            w.i32(remap[keys.index(('ax', lf.mdim))] if lf.mode else 0)
            for x in st:
                w.i64(x)
        self.codeobj.write(w)


class ExprBwdOp(Op):
    code = OP_EXPR_BWD

    def __init__(self, gleaf, fwd: ExprOp, target, kept, loop, gout, nsplit=1, acc=1, scale=1.0):
        self.gleaf, self.fwd, self.target, self.kept, self.loop = gleaf, fwd, target, kept, loop
        self.gout, self.nsplit, self.acc, self.scale = gout, nsplit, acc, scale

    def payload(self, w):
        w.tref(self.gleaf); w.i32(self.acc); w.f64(self.scale); w.i32(self.target); w.i32(self.nsplit)
        dims = self.kept + self.loop
        leaves = self.fwd.codeobj.leaves
        keys = [(d[0], d[1]) for d in dims]
        pinned = {keys.index(('ax', lf.mdim)) for lf in leaves if lf.mode}
        strides = [_strides_like(self.gout, self.fwd.keep, dims)] + [[lf.stride(d) for d in dims] for lf in leaves]
        sizes, n_a, strides, remap = _coalesce([d[2] for d in dims], len(self.kept), strides, pinned)
        _write_dims(w, sizes, n_a)
        w.tref(self.gout)
        for x in strides[0]:
            w.i64(x)
        w.i32(len(leaves))
        for lf, st in zip(leaves, strides[1:]):
            w.tref(lf.pt); w.i32(lf.mode)
            w.i32(remap[keys.index(('ax', lf.mdim))] if lf.mode else 0)
            for x in st:
                w.i64(x)
        self.fwd.codeobj.write(w)


class ReduceOp(Op):
    """out[od] (+)= scale * R_{rd}( sum_f coeff_f F_f ) + cadd     (csrc/kernels.cuh reduce_*)"""
    code = OP_REDUCE

    def __init__(self, mode, out, od, rd, factors, acc=0, scale=1.0, cadd=0.0, nsplit=1, lse=None, gout=None,
                 lse_dims=None, gout_dims=None, tag='', thread_hint=None):
        self.mode, self.out, self.od, self.rd, self.factors = mode, out, od, rd, factors
        self.thread_hint = _outputs_contiguous(factors, od, rd) if thread_hint is None else thread_hint
        self.acc, self.scale, self.cadd, self.nsplit = acc, scale, cadd, nsplit
        self.lse, self.gout, self.lse_dims, self.gout_dims, self.tag = lse, gout, lse_dims, gout_dims, tag
        # LSE modes whose output is differentiated also store the pair (m, lo) = (max, log(sum exp(s - m) + eps)): the
        # adjoint forms the softmax weight as exp((s - m) - lo), like autograd does for the reference, instead of
        # exp(s - lse) with lse = m + lo ROUNDED to the working precision -- at |lse| ~ 4e5 (cfg-5's top level) that
        # rounding alone is 3e-2 absolute, a 1.5 % error common to every weight.  WSUM: `lse` = (m PT, lo PT).
        self.m_out = self.lo_out = None

    def payload(self, w):
        w.i32(self.mode); w.tref(self.out); w.i32(self.acc); w.f64(self.scale); w.f64(self.cadd); w.i32(self.nsplit)
        w.i32(1 if self.thread_hint else 0)
        if self.mode in (R_LSE_EPS, R_LSE):
            w.i32(1 if self.m_out is not None else 0)
            if self.m_out is not None:
                w.tref(self.m_out); w.tref(self.lo_out)
        dims = self.od + self.rd
        if len(self.factors) > MAXL:
            raise Exception("contraction step joins more than 10 factor tensors")
        strides = [[lf.stride(d) for d in dims] for lf, _ in self.factors]
        extra = []
        if self.mode == R_WSUM:
            m_pt, lo_pt = self.lse
            extra = [(m_pt, self.lse_dims), (lo_pt, self.lse_dims), (self.gout, self.gout_dims)]
            strides += [_strides_like(pt, own, dims) for pt, own in extra]
        sizes, n_a, strides, _ = _coalesce([d[2] for d in dims], len(self.od), strides)
        _write_dims(w, sizes, n_a)
        w.i32(len(self.factors))
        for (lf, coeff), st in zip(self.factors, strides):
            w.f64(coeff)
            w.tref(lf.pt)
            for x in st:
                w.i64(x)
        for (pt, own), st in zip(extra, strides[len(self.factors):]):
            w.tref(pt)
            for x in st:
                w.i64(x)


class NormalFanOp(Op):
    """out[rows, f] = sum_d log N(v[rows,d]; l[rows,d], s[f,d])   (csrc/fused.cuh normal_fan_kernel)"""
    code = OP_NORMAL_FAN

    def __init__(self, out, D, rows, v, l, s, fan_axis, F, tag=''):
        self.out, self.D, self.rows, self.v, self.l, self.s = out, D, rows, v, l, s
        self.fan_axis, self.F, self.tag = fan_axis, F, tag

    def payload(self, w):
        w.tref(self.out); w.i32(self.D); w.i32(len(self.rows))
        for d in self.rows:
            w.i32(d[2])
        o = plain(self.out)
        for lf in (self.v, self.l, o):
            for d in self.rows:
                w.i64(lf.stride(d))
        ev = ('ev', 0, self.D)
        w.tref(self.v.pt); w.i64(self.v.stride(ev))
        w.tref(self.l.pt); w.i64(self.l.stride(ev))
        fdim = ('ax', self.fan_axis, self.F) if self.fan_axis else None
        w.tref(self.s.pt); w.i64(self.s.stride(fdim) if fdim else 0); w.i64(self.s.stride(ev))
        w.i32(self.F)
        w.i64(o.stride(fdim) if fdim else 0)


class NormalFanBwdOp(Op):
    """Adjoint of NormalFanOp (csrc/fused.cuh fan_bwd_rows_kernel / fan_bwd_scale_kernel).
    which = 0: R[rows, d] = 2 (v - l) sum_f G[rows, f] / (2 scale[f,d]^2);
    which = 1: per-CTA partials of V[f, d] = sum_rows G (v - l)^2 and Wsum[f] = sum_rows G."""
    code = OP_FAN_BWD

    def __init__(self, which, fan, gout, R, partial, partial_w, n_cta):
        self.which, self.fan, self.gout, self.R, self.partial, self.partial_w, self.n_cta = \
            which, fan, gout, R, partial, partial_w, n_cta

    def payload(self, w):
        f = self.fan
        w.i32(self.which); w.tref(self.gout); w.tref(self.R); w.tref(self.partial); w.tref(self.partial_w)
        w.i32(self.n_cta)
        w.i32(f.D); w.i32(len(f.rows))
        for d in f.rows:
            w.i32(d[2])
        o = plain(f.out)
        for lf in (f.v, f.l, o):
            for d in f.rows:
                w.i64(lf.stride(d))
        ev = ('ev', 0, f.D)
        w.tref(f.v.pt); w.i64(f.v.stride(ev))
        w.tref(f.l.pt); w.i64(f.l.stride(ev))
        fdim = ('ax', f.fan_axis, f.F) if f.fan_axis else None
        w.tref(f.s.pt); w.i64(f.s.stride(fdim) if fdim else 0); w.i64(f.s.stride(ev))
        w.i32(f.F)
        w.i64(o.stride(fdim) if fdim else 0)


class FanLseOp(Op):
    """out[rho, f] = LSE_eps over kappa of (Normal factor + small factors) + cadd, nothing materialised
    (csrc/fused.cuh fan_lse_kernel).  gen_expr / gen_reduce are the unfused ops it stands for."""
    code = OP_FAN_LSE

    def __init__(self, out, D, rho, kappa, v, l, s, fan_axis, F, bfactors, cadd, gen_expr, gen_reduce, tag=''):
        self.out, self.D, self.rho, self.kappa, self.v, self.l, self.s = out, D, rho, kappa, v, l, s
        self.fan_axis, self.F, self.bfactors, self.cadd = fan_axis, F, bfactors, cadd
        self.gen_expr, self.gen_reduce, self.tag = gen_expr, gen_reduce, tag
        self.dense = None            # (lam dim, L, NG) when the planner commits the adjoint to the dense kernel's gS layout
        self.psum = None             # (partial PT, rows, od) when the dense kernel also sums its output over the users
        self.qterm = None            # (ExprOp E, loc leaf, scale leaf, coeff): Gaussian Q factor evaluated inside the dense kernel

    def _body(self, w):
        w.i32(self.D); w.i32(len(self.rho))
        for d in self.rho:
            w.i32(d[2])
        o = plain(self.out)
        for lf in (self.v, self.l, o):
            for d in self.rho:
                w.i64(lf.stride(d))
        w.i32(self.kappa[2]); w.i64(self.v.stride(self.kappa)); w.i64(self.l.stride(self.kappa))
        ev = ('ev', 0, self.D)
        fdim = ('ax', self.fan_axis, self.F)
        w.tref(self.v.pt); w.i64(self.v.stride(ev))
        w.tref(self.l.pt); w.i64(self.l.stride(ev))
        w.tref(self.s.pt); w.i64(self.s.stride(fdim)); w.i64(self.s.stride(ev))
        w.i32(self.F); w.i64(o.stride(fdim))
        if len(self.bfactors) > MAXL:
            raise Exception("fused contraction joins more than 10 small factors")
        w.i32(len(self.bfactors))
        for lf, coeff in self.bfactors:
            w.f64(coeff); w.tref(lf.pt)
            for d in self.rho:
                w.i64(lf.stride(d))
            w.i64(lf.stride(self.kappa))
        w.f64(self.cadd)
        if self.qterm is None:
            w.i32(0)
        else:
            _, ql, qs, qc = self.qterm
            w.i32(1); w.tref(ql.pt); w.tref(qs.pt)
            for lf in (ql, qs):
                for d in self.rho:
                    w.i64(lf.stride(d))
            w.i64(ql.stride(ev)); w.i64(qs.stride(ev)); w.f64(qc)

    def payload(self, w):
        w.i32(0); w.tref(self.out)
        self._body(w)
        # fused plate sum (csrc/fan_tc2.cuh only): per-(CTA, team) partial sums over the users, [rows, od...]
        if self.psum is None:
            w.i32(0)
        else:
            part, rows, od = self.psum
            ref = _PartialRef(part, od, rows)
            w.i32(rows); w.tref(part)
            w.i64(ref.stride(self.dense[0])); w.i64(ref.stride(('ax', self.fan_axis, self.F))); w.i64(ref.stride(('sp', 0, rows)))


DENSE_TILES, DENSE_MAXG, DENSE_EVENTS = 3, 8, (2, 4, 6, 8, 12, 16, 18)
FAN_PSUM_ROWS = 160          # partial rows of the fused plate sum: one per CTA of the dense kernel (148 on B200)


def dense_fan_geometry(op: 'FanLseOp', itemsize=4):
    """Mirror of csrc/fan_tc2.cuh fan_lse_tc2_supported: (lam dim, L, number of fan groups) when the dense tcgen05
    formulation applies to this fused contraction (fp32, loc depends on exactly one rho dim that neither the value
    nor the small factors carry), else None.  Used by bench.py to describe the kernel that ran."""
    if itemsize != 4 or op.D not in DENSE_EVENTS or op.l.stride(op.kappa) != 0:
        return None
    lam = [d for d in op.rho if d[2] > 1 and op.l.stride(d) != 0]
    if len(lam) != 1:
        return None
    lam = lam[0]
    if op.v.stride(lam) != 0 or any(lf.stride(lam) != 0 for lf, _ in op.bfactors):
        return None
    L = lam[2]
    FP = L * op.F
    NG = -(-FP // (DENSE_TILES * 128))
    n_u = _prod(d[2] for d in op.rho) // L
    if FP < 96 or NG > DENSE_MAXG or NG > L or op.kappa[2] > 32 or n_u < 16 or len(op.bfactors) > 4 or len(op.rho) - 1 > 4:
        return None
    return lam, L, NG


class FanLseBwdOp(Op):
    """gS[rho, kappa] = sum_f gout[rho,f] * softmax weight: adjoint of the small-factor sum."""
    code = OP_FAN_LSE

    def __init__(self, fwd: FanLseOp, gout, gS, gout_dims=None):
        # gout_dims: dims the gout tensor is laid out over (default: like fwd.out); a plate-sum adjoint is
        # read in place through zero strides instead of being broadcast into a full-size tensor first
        self.fwd, self.gout, self.gS, self.gout_dims = fwd, gout, gS, gout_dims

    def payload(self, w):
        w.i32(1); w.tref(self.fwd.out); w.tref(self.gout); w.tref(self.gS)
        self.fwd._body(w)
        f = self.fwd
        fdim = ('ax', f.fan_axis, f.F)
        if self.gout_dims is None:
            o = plain(f.out)
            st = [o.stride(d) for d in f.rho] + [o.stride(fdim)]
        else:
            st = _strides_like(self.gout, self.gout_dims, f.rho + [fdim])
        for x in st:
            w.i64(x)
        # > 0: gS is laid out [users, NG fan-group partials, kappa] (csrc/fan_tc2.cuh only); 0: [rho, kappa]
        w.i32(self.fwd.dense[2] if self.fwd.dense is not None else 0)


class NormalQBwdOp(Op):
    """Whole adjoint of a mean-field Gaussian Q factor in one pass over its value (csrc/qfactor.cuh):
    G[u, kappa] = coeff * sum_s gS[u, s, kappa];  g_loc += sum_kappa G (v - loc) / scale^2;
    g_ls += sum_kappa G ((v - loc)^2 / scale^2 - 1)  (or the gradient w.r.t. the scale itself).
    `gen` = what the generic path would have done (the emulator runs that)."""
    code = OP_NORMAL_Q_BWD

    def __init__(self, D, users, kappa, v, l, s, scale_is_exp, gS, gS_dims, sdim, coeff, g_l, g_s):
        self.D, self.users, self.kappa, self.v, self.l, self.s = D, users, kappa, v, l, s
        self.scale_is_exp, self.gS, self.gS_dims, self.sdim, self.coeff = scale_is_exp, gS, gS_dims, sdim, coeff
        self.g_l, self.g_s, self.acc_l, self.acc_s = g_l, g_s, 1, 1

    def payload(self, w):
        w.i32(self.D); w.i32(len(self.users))
        for d in self.users:
            w.i32(d[2])
        gref = _OwnDims(self.gS, self.gS_dims)
        for lf in (self.v, self.l, self.s, gref):
            for d in self.users:
                w.i64(lf.stride(d))
        ev = ('ev', 0, self.D)
        w.i32(self.kappa[2]); w.i64(self.v.stride(self.kappa)); w.i64(self.v.stride(ev))
        w.i64(self.l.stride(ev)); w.i64(self.s.stride(ev))
        w.tref(self.v.pt); w.tref(self.l.pt); w.tref(self.s.pt); w.i32(1 if self.scale_is_exp else 0)
        w.tref(self.gS); w.i64(gref.stride(self.sdim) if self.sdim is not None else 0); w.i64(gref.stride(self.kappa))
        w.i32(self.sdim[2] if self.sdim is not None else 1)
        w.f64(self.coeff)
        for g, a in ((self.g_l, self.acc_l), (self.g_s, self.acc_s)):
            w.i32(1 if g is not None else 0)
            if g is not None:
                w.tref(g); w.i32(a)


class PermOp(Op):
    """perm[rows, K] (int64) from float64 uniforms u[rows, K]: mode 0 = argsort along K (PermutationSampler.perm,
    reference Sampler.py:143-148), mode 1 = floor(u K) (CategoricalSampler.perm).  csrc/sampling.cuh perm_kernel."""
    code = OP_PERM

    def __init__(self, u, out, rows, K, mode):
        self.u, self.out, self.rows, self.K, self.mode = u, out, rows, K, mode

    def payload(self, w):
        w.tref(self.u); w.tref(self.out); w.i64(self.rows); w.i32(self.K); w.i32(self.mode)


class KGatherOp(Op):
    """out[o, k, i] = x[o, perm[o, k], i]: parent particles permuted along their K axis (Sampler.resample_scope,
    reference Sampler.py:85-116).  csrc/sampling.cuh kgather_kernel."""
    code = OP_KGATHER

    def __init__(self, x, perm, out, outer, K, inner):
        self.x, self.perm, self.out, self.outer, self.K, self.inner = x, perm, out, outer, K, inner

    def payload(self, w):
        w.tref(self.x); w.tref(self.perm); w.tref(self.out); w.i64(self.outer); w.i64(self.K); w.i64(self.inner)


class TsSampleOp(Op):
    """The T-step recursion of a Timeseries draw in one launch (reference Timeseries.py:89-123): `expr` (an ExprOp,
    nothing summed, dims [outer..., T, K, event...]) is evaluated step by step with leaf `prev_leaf` = the previous
    step's draw, permuted along K by perm[outer, t, :] (timeseries_perm).  csrc/sampling.cuh ts_sample_kernel."""
    code = OP_TS_SAMPLE

    def __init__(self, expr: 'ExprOp', prev_leaf, t_dim, k_dim, init, perm, n_outer, T, K, E):
        self.expr, self.prev_leaf, self.t_dim, self.k_dim = expr, prev_leaf, t_dim, k_dim
        self.init, self.perm, self.n_outer, self.T, self.K, self.E = init, perm, n_outer, T, K, E
        self.out = expr.out

    def payload(self, w):
        e = self.expr
        w.tref(e.out)
        dims = e.keep
        leaves = e.codeobj.leaves
        _write_dims(w, [d[2] for d in dims], len(dims))
        w.i32(len(leaves))
        for lf in leaves:
            w.tref(lf.pt); w.i32(lf.mode); w.i32(0)
            for d in dims:
                w.i64(lf.stride(d))
        e.codeobj.write(w)
        w.i32(self.prev_leaf); w.i32(self.t_dim); w.i32(self.k_dim)
        w.tref(self.init)
        w.i32(1 if self.perm is not None else 0)
        if self.perm is not None:
            w.tref(self.perm)
        w.i64(self.n_outer); w.i32(self.T); w.i32(self.K); w.i32(self.E)


class PasteOp(Op):
    """dst[i0, i1, ...] = src[i0, i1, ...] for i_k < size_k: a block of a smaller tensor written into the leading corner
    of a larger one (prediction: the posterior sample / the observed data pasted into the extended draw, reference
    dist.py:256-267).  dims: [(size, src stride, dst stride)].  csrc/sampling.cuh paste_kernel."""
    code = OP_PASTE

    def __init__(self, src, dst, dims):
        self.src, self.dst, self.dims, self.out = src, dst, [tuple(int(x) for x in d) for d in dims], dst

    def payload(self, w):
        w.tref(self.src); w.tref(self.dst)
        dims = [d for d in self.dims if d[0] > 1] or [(1, 0, 0)]
        if len(dims) > MAXD:
            raise Exception("paste: too many dims")
        w.i32(len(dims))
        for size, ss, ds in dims:
            w.i32(size); w.i64(ss); w.i64(ds)


class ReduceSeqOp(Op):
    """A run of consecutive SMALL reductions (top-level contractions, their adjoints, the cross-rank sum between them)
    executed by ONE single-CTA launch (csrc/kernels.cuh reduce_seq_kernel): the ops run back to back with a
    __syncthreads() between them instead of a launch latency + drain each (~4 us per dependent graph node for
    nanoseconds of work).  Built by `fuse_small_reductions` after the programs are complete; to the dependency
    analysis it is one op touching the union of its members' tensors."""
    code = OP_RSEQ
    MAX_OPS, MAX_POINTS = 12, 16384
    # Measured on B200 (profiles/r02_small_ops.md): a member costs ~4.5 us inside the single CTA (a serial chain of ~2000
    # dependent instructions per warp task, IPC 0.25: ncu), about what a dependent node of a replayed CUDA graph costs,
    # and separate nodes overlap with the rest of the program's DAG.  Runs of 4-5 members (MovieLens top level) LOSE
    # 14 us per step, runs of 7-9 (radon-shaped trees) are neutral (0.251 vs 0.256 ms per marginals() replay, 34 vs 48
    # launches): only long runs are fused.
    MIN_OPS = int(os.environ.get('ALAN_B200_RSEQ_MIN', '6'))

    def __init__(self, ops):
        self.ops = list(ops)
        self.is_barrier = any(isinstance(o, XReduceOp) for o in self.ops)
        self.tag = '+'.join(getattr(o, 'tag', '') or type(o).__name__ for o in self.ops)

    def payload(self, w):
        w.i32(len(self.ops))
        for o in self.ops:
            o.serialize(w)


def fuse_small_reductions(prog):
    """Wrap maximal runs of consecutive small ReduceOps (and XReduceOps between / beside them) into ReduceSeqOps."""
    def small(o):
        if type(o) is XReduceOp:
            return True
        if type(o) is not ReduceOp:
            return False
        pts = _prod(d[2] for d in o.od + o.rd)
        return pts * max(len(o.factors), 1) <= ReduceSeqOp.MAX_POINTS
    out, run = [], []

    def close():
        nonlocal run
        n_red = sum(type(o) is ReduceOp for o in run)
        n_x = sum(type(o) is XReduceOp for o in run)
        if n_red + n_x >= max(ReduceSeqOp.MIN_OPS, 2):
            out.append(ReduceSeqOp(run))
        else:
            out.extend(run)
        run = []
    for o in prog:
        if small(o) and len(run) < ReduceSeqOp.MAX_OPS and not (type(o) is XReduceOp and any(type(x) is XReduceOp for x in run)):
            run.append(o)
        else:
            close()
            if small(o):
                run.append(o)
            else:
                out.append(o)
    close()
    return out


class MvnPrepOp(Op):
    """csrc/mvn.cuh: the matrix argument of a MultivariateNormal -> scale_tril L, its inverse W and the constant
    c = -sum log L_ii - d/2 log 2 pi, one warp per matrix.  mode: 0 covariance_matrix, 1 precision_matrix, 2 scale_tril."""
    code = OP_MVN_PREP
    MODES = {'covariance_matrix': 0, 'precision_matrix': 1, 'scale_tril': 2}

    def __init__(self, S, L, W, c, n_mat, d, mode, low_rank=None):
        self.S, self.L, self.W, self.c, self.n_mat, self.d, self.mode = S, L, W, c, n_mat, d, mode
        # mode 3 (LowRankMultivariateNormal): S is the factor [n_mat, d, r] and low_rank = (diagonal PT [n_mat, d], r);
        # the covariance S S^T + diag(.) is formed in shared memory and factorised like mode 0
        self.low_rank = low_rank
        self.out = W

    def payload(self, w):
        for pt in (self.S, self.L, self.W, self.c):
            w.tref(pt)
        w.i64(self.n_mat); w.i32(self.d); w.i32(self.mode)
        w.i32(self.low_rank[1] if self.low_rank else 0)
        if self.low_rank:
            w.tref(self.low_rank[0])


class DotOp(Op):
    """out[keep] = sum_e a * b   (csrc/fused.cuh dot_kernel)"""
    code = OP_DOT

    def __init__(self, out, keep, red, a, b, tag=''):
        self.out, self.keep, self.red, self.a, self.b, self.tag = out, keep, red, a, b, tag

    def payload(self, w):
        w.tref(self.out)
        dims = self.keep + self.red
        strides = [[lf.stride(d) for d in dims] for lf in (self.a, self.b)]
        sizes, n_a, strides, _ = _coalesce([d[2] for d in dims], len(self.keep), strides)
        _write_dims(w, sizes, n_a)
        for lf, st in zip((self.a, self.b), strides):
            w.tref(lf.pt)
            for x in st:
                w.i64(x)


class BernDotSumOp(Op):
    """out[od] = cadd + sum_rd log Bernoulli(y; logits = sum_e a * b)   (csrc/fused.cuh bern_dot_sum_kernel).
    `gen_ops` = the unfused (dot, expression, plate sum) ops it stands for."""
    code = OP_BERN_DOT

    def __init__(self, out, od, rd, D, a, b, y, cadd, gen_ops, tag=''):
        self.out, self.od, self.rd, self.D, self.a, self.b, self.y = out, od, rd, D, a, b, y
        self.cadd, self.gen_ops, self.tag = cadd, gen_ops, tag
        # (ExprOp E, loc leaf, scale leaf): `E.out[od] = sum_e log N(a; loc, scale)`, a Gaussian factor of the SAME rows
        # `a` that the kernel holds in registers anyway, written as a second output (Planner.fuse_side_factors)
        self.side = None

    def payload(self, w):
        w.tref(self.out); w.f64(self.cadd); w.i32(self.D)
        dims = self.od + self.rd
        leaves = [self.a, self.b, self.y] + ([self.side[1], self.side[2]] if self.side is not None else [])
        strides = [[lf.stride(d) for d in dims] for lf in leaves]
        sizes, n_a, strides, _ = _coalesce([d[2] for d in dims], len(self.od), strides)
        _write_dims(w, sizes, n_a)
        ev = ('ev', 0, self.D)

        def opnd(lf, st, has_ev):
            w.tref(lf.pt)
            for x in st:
                w.i64(x)
            if has_ev:
                w.i64(lf.stride(ev))
        opnd(self.a, strides[0], True); opnd(self.b, strides[1], True); opnd(self.y, strides[2], False)
        w.i32(0 if self.side is None else 1)
        if self.side is not None:
            w.tref(self.side[0].out)
            opnd(self.side[1], strides[3], True); opnd(self.side[2], strides[4], True)


class NormalPolySumOp(Op):
    """out[rows, K...] = cadd + sum_z log N(value; loc, scale) with `value - loc` a polynomial whose monomials are
    (Z leaves: tensors over the summed plate) x (K leaves: tensors over K axes), scale a K leaf or a constant
    (csrc/normal_poly.cuh).  `zterms` / `kterms`: [(coeff, [z leaf ids], [k leaf ids])]; `gen_ops` = the unfused
    (expression, plate sum) ops it stands for."""
    code = OP_NORMAL_POLY_SUM

    def __init__(self, out, rows, kd, zd, zleaves, kleaves, zterms, kterms, scale_leaf, scale_const, cadd, gen_ops, tag=''):
        self.out, self.rows, self.kd, self.zd = out, rows, kd, zd
        self.zleaves, self.kleaves, self.zterms, self.kterms = zleaves, kleaves, zterms, kterms
        self.scale_leaf, self.scale_const, self.cadd, self.gen_ops, self.tag = scale_leaf, scale_const, cadd, gen_ops, tag

    def payload(self, w):
        dims = self.rows + self.kd + self.zd
        w.tref(self.out); w.f64(self.cadd)
        w.i32(len(self.rows)); w.i32(len(self.kd)); w.i32(len(self.zd))
        for d in dims:
            w.i32(d[2])
        o = plain(self.out)
        for d in self.rows + self.kd:
            w.i64(o.stride(d))
        for leaves in (self.zleaves, self.kleaves):
            w.i32(len(leaves))
            for lf in leaves:
                w.tref(lf.pt)
                for d in dims:
                    w.i64(lf.stride(d))
        for terms in (self.zterms, self.kterms):
            w.i32(len(terms))
            for coeff, zs, ks in terms:
                w.f64(coeff)
                for ids in (zs, ks):
                    ids = list(ids) + [-1, -1]
                    w.i32(ids[0]); w.i32(ids[1])
        w.i32(self.scale_leaf); w.f64(self.scale_const)


class ChainOp(Op):
    code = OP_CHAIN

    def __init__(self, ms, levels, out, outer, T, K):
        self.ms, self.levels, self.out, self.outer, self.T, self.K = ms, levels, out, outer, T, K

    def payload(self, w):
        w.tref(self.ms); w.tref(self.levels); w.tref(self.out)
        w.i64(self.outer); w.i64(self.T); w.i64(self.K)


class ChainBwdOp(Op):
    code = OP_CHAIN_BWD

    def __init__(self, fwd: ChainOp, gout, glevels, gms):
        self.fwd, self.gout, self.glevels, self.gms = fwd, gout, glevels, gms

    def payload(self, w):
        f = self.fwd
        w.tref(f.ms); w.tref(f.levels); w.tref(f.out); w.tref(self.gout); w.tref(self.glevels); w.tref(self.gms)
        w.i64(f.outer); w.i64(f.T); w.i64(f.K)


class SampleOp(Op):
    code = OP_SAMPLE

    def __init__(self, batch, ks, factors, idx_tensors, u, outs):
        # batch: dims; ks: dims; factors: [(LeafRef, coeff, gathered [(axis, idx slot)])];
        # idx_tensors: [(PT, own dims)]; u: (PT, own dims); outs: [PT]
        self.batch, self.ks, self.factors, self.idx_tensors, self.u, self.outs = batch, ks, factors, idx_tensors, u, outs

    def payload(self, w):
        if len(self.batch) > MAXD or len(self.ks) > 4 or len(self.idx_tensors) > 16:
            raise Exception("resampling step exceeds kernel limits")
        w.i32(len(self.batch))
        for d in self.batch:
            w.i32(d[2])
        w.i32(len(self.ks))
        for d in self.ks:
            w.i32(d[2])
        w.i32(len(self.factors))
        for (lf, coeff, gathered) in self.factors:
            w.f64(coeff)
            _opnd(w, lf, self.batch, False)
            for d in self.ks:
                w.i64(lf.stride(d))
            if len(gathered) > 6:
                raise Exception("factor depends on more than 6 already-sampled K axes")
            w.i32(len(gathered))
            for (axis, slot) in gathered:
                w.i64(lf.stride(('ax', axis, 0)))
                w.i32(slot)
        w.i32(len(self.idx_tensors))
        for (pt, own) in self.idx_tensors:
            w.tref(pt)
            for s in _strides_like(pt, own, self.batch):
                w.i64(s)
        pt, own = self.u
        w.tref(pt)
        for s in _strides_like(pt, own, self.batch):
            w.i64(s)
        for o in self.outs:
            w.tref(o)


# ----------------------------------------------------------------------------------------
# signature of the call
# ----------------------------------------------------------------------------------------
@dataclass
class TensorSig:
    role: str                # 'sample' | 'param' | 'data' | 'elf'
    axes: tuple
    pos_shape: tuple
    requires_grad: bool = False


@dataclass
class LogicalFactor:
    tensors: list            # [(LeafRef, coeff)]
    const: float
    axes: tuple


@dataclass
class Step:
    level: tuple             # active plates at this level
    tensors: list            # [(LeafRef, coeff)]
    ks: tuple                # K axes sampled jointly at this step (row-major order)


class Plan:
    def __init__(self):
        self.dtype = None
        self.input_names = []        # order of device pointers handed to the ABI
        self.input_pts = {}
        self.const_inputs = {}       # name -> torch tensor (uploaded once by the runtime)
        self.programs = []           # list of op lists: fwd segments, bwd segments, sample
        self.n_fwd = self.n_bwd = 0
        self.sample_prog = -1
        self.ws_bytes = 0
        self.grad_inputs = []        # names of inputs whose gradient the bwd program writes
        self.sample_steps = []       # [(batch axes (without N), ks)] in visiting order
        self.sample_groups = []      # [(groupvarname, plate axes)] = idx outputs
        self.sizes = {}
        self.N = None
        self.canon_axes = None
        self.allreduce = None        # (PT of the tile, n elements) for plate sharding
        self.fused_collectives = False   # True: the cross-rank sums are XReduceOps inside the programs
        self.global_grads = []       # grads that must be all-reduced across shards
        self.blob = None
        self.retained = {}           # debugging: name -> PT of interesting intermediates

    def insert_deps(self):
        """Open every program with its dependency table (DepsOp).  Two ops are ordered when they touch the same
        workspace or output tensor (inputs and aux tensors are read-only; which of the two writes is not tracked, so two
        readers of one intermediate stay ordered too -- conservative); ops whose footprint is not a list of tensors
        (region fills, cross-rank reductions) are ordered against everything."""
        for pi, prog in enumerate(self.programs):
            ops = [op for op in prog if not isinstance(op, DepsOp)]
            if len(ops) <= 2:
                self.programs[pi] = ops
                continue
            last = {}                      # tensor key -> index (in the final program, DepsOp = 0) of its last toucher
            deps, barrier = [[]], 0        # entry 0: the table itself
            for k, op in enumerate(ops, start=1):
                if isinstance(op, (FillRegionOp, XReduceOp)) or getattr(op, 'is_barrier', False):
                    deps.append(list(range(1, k)))
                    barrier = k
                    last = {}
                    continue
                w = W()
                op.payload(w)
                keys = {(pt.space, pt.id if pt.space == 'ws' else pt.index) for pt in w.refs if pt.space in ('ws', 'output')}
                d = {last[key] for key in keys if key in last}
                if barrier:
                    d.add(barrier)
                deps.append(sorted(d))
                for key in keys:
                    last[key] = k
            self.programs[pi] = [DepsOp(deps)] + ops

    def assign_offsets(self, itemsize):
        """Dry-run serialisation to find the workspace tensors the emitted ops really touch, then lay
        them out: forward tensors first, adjoints/partials after (one contiguous region to zero)."""
        self.adj_region = (0, 0)
        if os.environ.get('ALAN_B200_RSEQ', '1') != '0':       # read when the plan is BUILT
            self.programs = [fuse_small_reductions([o for o in prog if not isinstance(o, DepsOp)]) for prog in self.programs]
        self.insert_deps()
        w = W()
        for prog in self.programs:
            for op in prog:
                op.serialize(w)
        off = 0
        for group in (0, 1):
            if group == 1:
                lo = off
            for pt in sorted(w.seen.values(), key=lambda p: p.id):
                if getattr(pt, 'group', 0) != group:
                    continue
                pt.offset = off
                nbytes = getattr(pt, 'alloc_numel', pt.numel) * itemsize
                off += (nbytes + 255) // 256 * 256
        self.adj_region = (lo, off)
        self.ws_bytes = max(off, 256)

    def serialize(self):
        w = W()
        for v in (MAGIC, VERSION, 0 if self.dtype == torch.float32 else 1, len(self.input_names),
                  len(self.programs)):
            w.i32(v)
        w.i64(self.ws_bytes)
        w.i32(self.n_fwd); w.i32(self.n_bwd); w.i32(self.sample_prog)
        table = len(w.w)
        for _ in self.programs:
            w.i32(0); w.i32(0)
        for i, prog in enumerate(self.programs):
            w.w[table + 2 * i] = len(w.w)
            w.w[table + 2 * i + 1] = len(prog)
            for op in prog:
                op.serialize(w)
        self.blob = torch.tensor(w.w, dtype=torch.int32)
        return self.blob


# ----------------------------------------------------------------------------------------
# the planner
# ----------------------------------------------------------------------------------------
class Planner:
    def __init__(self, P: Plate, Q: Plate, sig: dict, sizes: dict, dtype, extra_factors=(), want_sample_N=None,
                 shard_plate=None, world_size=1, constants=None, fast_paths=True, grad_names=(),
                 fused_collectives=False, nonmp=False):
        """sig: name -> TensorSig for samples, inputs/params, data and tensor-valued extra factors.
        sizes: axis name -> extent (plates and K axes).
        extra_factors: [(key, Expr)] expressions over input leaves, added as log factors at the plate
        level matching their plate axes (reference Sample.py:69-108 `extra_log_factors`)."""
        self.P, self.Q, self.sig, self.sizes, self.dtype = P, Q, sig, dict(sizes), dtype
        self.itemsize = 4 if dtype == torch.float32 else 8
        self.extra_factors = list(extra_factors)
        self.N = want_sample_N
        self.fast_paths = fast_paths
        self.shard_plate, self.world_size = shard_plate, world_size
        # fused_collectives: the cross-rank sums are ops INSIDE the programs (XReduceOp over symmetric memory) instead
        # of host-issued all-reduces between program segments: a sharded plan then has one forward and one backward
        # program like an unsharded one
        self.fused_collectives = bool(fused_collectives) and shard_plate is not None
        self.plan = Plan()
        self.plan.dtype = dtype
        self.plan.sizes = self.sizes
        self.plan.fused_collectives = self.fused_collectives
        self.all_plates = P.all_platenames()
        self.groups = Q.groupvarnames()
        self.v2g = Q.varname2groupvarname()
        self.g2plates = Q.groupvarname2platenames()
        # nonmp: every latent shares ONE K axis (reference SampleNonMP.py:22-26 `unify_dims`), no K is contracted
        self.nonmp = bool(nonmp)
        self.canon = list(self.all_plates) + ([NONMP_K] if self.nonmp else [Kname(g) for g in self.groups])
        self.plan.canon_axes = self.canon
        self.ws_off = 0
        self.alloc_group = 0
        self.grad_names = list(grad_names)
        self.needs = set()
        self.needs_materialised = set()      # factor tensors someone reads outside the contraction that produced them
        self.fan_by_out = {}
        self.producer = {}
        self.fwd = []
        self.fwd_segments = []
        self.steps = []              # resampling steps in forward (bottom-up) creation order per level
        self.level_steps = {}        # level tuple -> [Step]
        self.level_factors = {}      # level tuple -> (factors before contraction, Ks summed there)
        self.level_paths = {}        # level tuple -> forward contraction path
        self.level_order = []
        self.consts = {}
        self.inputs = {}
        for name, s in sig.items():
            self._add_input(name, s.axes, s.pos_shape)
        self.scope = {}
        for name, s in sig.items():
            if s.role in ('sample', 'param'):
                self.scope[name] = Expr.leaf(self.inputs[name], s.axes, s.pos_shape)

    @classmethod
    def bare(cls, sig: dict, sizes: dict, dtype, canon):
        """A planner without a model tree: only inputs, workspace tensors and `emit_expr` / `emit` -- used for the
        stand-alone weighted sums of `Marginals.moments` (weighted_moment_plan below)."""
        self = cls.__new__(cls)
        self.P = self.Q = None
        self.sig, self.sizes, self.dtype = sig, dict(sizes), dtype
        self.itemsize = 4 if dtype == torch.float32 else 8
        self.extra_factors, self.N, self.fast_paths = [], None, False
        self.shard_plate, self.world_size = None, 1
        self.plan = Plan()
        self.plan.dtype, self.plan.sizes = dtype, self.sizes
        self.all_plates = [a for a in canon if not a.startswith('K_')]
        self.groups, self.v2g, self.g2plates = [], {}, {}
        self.canon = list(canon)
        self.plan.canon_axes = self.canon
        self.alloc_group = 0
        self.grad_names, self.needs, self.needs_materialised = [], set(), set()
        self.fan_by_out, self.producer, self.fwd, self.fwd_segments = {}, {}, [], []
        self.consts, self.inputs = {}, {}
        for name, s in sig.items():
            self._add_input(name, s.axes, s.pos_shape)
        return self

    def set_grad_names(self, names):
        self.grad_names = list(names)
        self.needs = set(self.inputs[n].id for n in self.grad_names)

    @staticmethod
    def op_inputs(op):
        op = getattr(op, 'autodiff_as', None) or op
        if isinstance(op, ExprOp):
            return [lf.pt for lf in op.codeobj.leaves]
        if isinstance(op, ReduceOp):
            return [lf.pt for lf, _ in op.factors]
        if isinstance(op, ChainOp):
            return [op.ms]
        if isinstance(op, FanLseOp):
            q = [op.qterm[1].pt, op.qterm[2].pt] if op.qterm is not None else []
            return [lf.pt for lf, _ in op.bfactors] + [op.v.pt, op.l.pt, op.s.pt] + q
        if isinstance(op, BernDotSumOp):
            return [op.a.pt, op.b.pt, op.y.pt] + ([op.side[1].pt, op.side[2].pt] if op.side is not None else [])
        if isinstance(op, NormalPolySumOp):
            return [lf.pt for lf in op.zleaves + op.kleaves]
        if isinstance(op, DotOp):
            return [op.a.pt, op.b.pt]
        return []

    def emit(self, op):
        """Append a forward op and propagate 'needs a gradient' to its output."""
        self.fwd.append(op)
        self.producer[op.out.id] = op
        if not isinstance(getattr(op, 'autodiff_as', None), str):
            if any(p.id in self.needs for p in self.op_inputs(op)):
                self.needs.add(op.out.id)
                if isinstance(op, ReduceOp) and op.mode in (R_LSE_EPS, R_LSE) and op.m_out is None:
                    op.m_out = self.ws(op.out.axes, op.out.pos_shape, name='lse_m')
                    op.lo_out = self.ws(op.out.axes, op.out.pos_shape, name='lse_lo')

    # -- allocation -------------------------------------------------------------------
    def _add_input(self, name, axes, pos_shape):
        for a in axes:
            if a not in self.sizes:
                raise Exception(f"{name}: axis {a} has no known size")
        axes_c = self.canon_order(axes)
        if tuple(axes_c) != tuple(axes):
            raise Exception(f"input {name} must be laid out with named axes in canonical order {axes_c}, got {axes}")
        pt = PT(axes, pos_shape, self.sizes, 'input', index=len(self.plan.input_names), name=name)
        self.plan.input_names.append(name)
        self.plan.input_pts[name] = pt
        self.inputs[name] = pt
        return pt

    def const_input(self, value: torch.Tensor):
        key = (tuple(value.shape), tuple(value.reshape(-1).tolist()))
        if key not in self.consts:
            name = f"__const{len(self.consts)}"
            pt = self._add_input(name, (), tuple(value.shape))
            self.plan.const_inputs[name] = value.to(self.dtype).contiguous()
            self.consts[key] = pt
        return self.consts[key]

    def ws(self, axes, pos_shape=(), name=''):
        pt = PT(axes, pos_shape, self.sizes, 'ws', offset=None, name=name)
        pt.group = self.alloc_group
        return pt

    def ws_raw(self, numel, name=''):
        pt = PT((), (max(int(numel), 1),), self.sizes, 'ws', offset=None, name=name)
        pt.group = self.alloc_group
        return pt

    def canon_order(self, axes):
        axes = list(axes)
        known = [a for a in self.canon if a in axes]
        extra = [a for a in axes if a not in self.canon]
        return tuple(known + extra)

    def axdim(self, a):
        return ('ax', a, int(self.sizes[a]))

    # -- expression lowering ------------------------------------------------------------
    def _numel(self, e: Expr):
        n = 1
        for a in e.axes:
            n *= self.sizes[a]
        for s in e.pos_shape:
            n *= s
        return n

    def materialize(self, e: Expr, tag='') -> Expr:
        """Evaluate `e` into a workspace tensor; returns a leaf Expr reading it."""
        if e.op == 'leaf' and e.mode == 0 and not e.rename:
            return e
        if e.op == 'sumlast':
            body = self._prepare(e.args[0])
            pt, op = self.emit_expr(body, nred=1, tag=tag or 'sumlast', append=False)
            a_b = body.args if body.op == 'mul' else ()
            if self.fast_paths and len(a_b) == 2 and all(x.op == 'leaf' and x.mode == 0 and not x.rename for x in a_b):
                dot = DotOp(pt, op.keep, op.red, plain(a_b[0].ref), plain(a_b[1].ref), tag='dot')
                dot.autodiff_as = op
                self.emit(dot)
            else:
                self.emit(op)
        else:
            body = self._prepare(e)
            pt = self.emit_expr(body, nred=0, tag=tag or 'expr')
        return Expr.leaf(pt, pt.axes, pt.pos_shape)

    def _prepare(self, e: Expr, space=None, memo=None) -> Expr:
        """Replace inner reductions (and cheap-to-hoist subtrees) by materialised leaves.  Shared subexpressions stay
        shared (one VM register, one materialised tensor), they are not copied per use."""
        if e.op in ('leaf', 'const'):
            return e
        memo = {} if memo is None else memo
        if id(e) in memo:
            return memo[id(e)][1]
        if e.op == 'sumlast':
            out = self.materialize(e)
        else:
            args = [self._prepare(a, space, memo) for a in e.args]
            out = Expr(e.op, args, e.axes, e.pos_shape)
        memo[id(e)] = (e, out)               # keeps `e` alive: ids are not reused while the memo exists
        return out

    def _hoist_args(self, args, space_numel):
        out = []
        for a in args:
            if a.op not in ('leaf', 'const') and self._numel(a) * HOIST_RATIO <= space_numel:
                a = self.materialize(a, tag='arg')
            out.append(a)
        return out

    def _codegen(self, body: Expr) -> Code:
        instrs, consts, leaves = [], [], []
        memo = {}
        leaf_slot = {}

        def reg():
            if len(instrs) >= NREG:
                raise Exception("traced expression needs more than 32 VM registers")
            return len(instrs)

        def go(e: Expr):
            if id(e) in memo:
                return memo[id(e)]
            if e.op == 'leaf':
                lf = LeafRef(e.ref, tuple(sorted(e.rename.items())), e.mode, e.mdim)
                if lf not in leaf_slot:
                    leaf_slot[lf] = len(leaves)
                    leaves.append(lf)
                    r = reg()
                    instrs.append((VOPS['load'], r, leaf_slot[lf], 0, 0, 0))
                    memo[('L', lf)] = r
                r = memo[('L', lf)]
            elif e.op == 'const':
                if e.value not in consts:
                    consts.append(e.value)
                r = reg()
                instrs.append((VOPS['const'], r, consts.index(e.value), 0, 0, 0))
            else:
                rs = [go(a) for a in e.args]
                while len(rs) < 4:
                    rs.append(0)
                r = reg()
                instrs.append((VOPS[e.op], r, rs[0], rs[1], rs[2], rs[3]))
            memo[id(e)] = r
            return r
        res = go(body)
        return Code(instrs, consts, res, leaves)

    def emit_expr(self, body: Expr, nred, tag='', out=None, acc=0, scale=1.0, append=True):
        """body: elementwise tree; nred = number of trailing positional dims summed ('all' = every one).
        Returns the output tensor (and the op itself when append=False)."""
        R = len(body.pos_shape)
        if nred == 'all':
            nred = R
        axes = self.canon_order(body.axes)
        keep = [self.axdim(a) for a in axes]
        ev = [('ev', R - 1 - i, int(body.pos_shape[i])) for i in range(R)]
        keep_ev, red_ev = ev[:R - nred], ev[R - nred:]
        code = self._codegen(body)
        if out is None:
            out = self.ws(axes, body.pos_shape[:R - nred], name=tag)
        op = ExprOp(out, keep + keep_ev, red_ev, code, acc=acc, scale=scale, tag=tag)
        if not append:
            return out, op
        self.emit(op)
        return out

    # -- distribution arguments ------------------------------------------------------------
    def resolve_arg(self, family, argname, v, scope):
        """dist.py:211-229: number / tensor / scope string / lambda -> Expr"""
        if isinstance(v, str):
            if v not in scope:
                raise Exception(f"{v} is not in scope")
            return scope[v]
        if isinstance(v, types.FunctionType):
            names = function_arguments(v)
            for n in names:
                if n not in scope:
                    raise Exception(f"{n} is not in scope")
            return trace_function(v, [scope[n] for n in names])
        if isinstance(v, torch.Tensor):
            if v.ndim == 0:
                return Expr.const(float(v))
            pt = self.const_input(v)
            return Expr.leaf(pt, (), pt.pos_shape)
        assert isinstance(v, numbers.Number)
        return Expr.const(float(v))

    def density(self, dist: Dist, value: Expr, scope, tag) -> PT:
        """One factor tensor: sum over every positional dim of log p(value; args)
        (TorchDimDist.py:157-162)."""
        if dist.family in ('MultivariateNormal', 'LowRankMultivariateNormal'):
            return self._mvn_density(dist, value, scope, tag)
        if dist.family == 'Dirichlet':
            return self._dirichlet_density(dist, value, scope, tag)
        if dist.family in COMPOSED:
            return self._composed_density(dist, value, scope, tag)
        args = {k: self.resolve_arg(dist.family, k, v, scope) for k, v in dist.args.items()}
        if dist.family in ('Bernoulli',):
            opname = 'Bernoulli_logits' if 'logits' in args else 'Bernoulli_probs'
            order = [args['logits'] if 'logits' in args else args['probs']]
        elif dist.family in ('NegativeBinomial', 'Binomial'):
            which = 'logits' if 'logits' in args else 'probs'
            opname = f'{dist.family}_{which}'
            order = [args['total_count'], args[which]]
        else:
            opname = dist.family
            order = [args[k] for k in DENSITY_ARGS[dist.family]]
        operands = [self._prepare(value)] + [self._prepare(a) for a in order]
        probe = Expr(opname, operands, *_union(operands))
        operands = [operands[0]] + self._hoist_args(operands[1:], self._numel(probe))
        body = Expr(opname, operands, *_union(operands))
        out, op = self.emit_expr(body, nred='all', tag=tag, append=False)
        fan = self._try_normal_fan(opname, operands, body, out, tag) if self.fast_paths else None
        if fan is not None:
            fan.autodiff_as = op            # adjoints are derived from the generic form of the same factor ...
            op.fan_twin = fan               # ... unless the fused adjoint kernels apply (build_backward)
            self.fan_by_out[out.id] = fan
            self.emit(fan)
        else:
            self.emit(op)
        return out

    # -- MultivariateNormal (dist.py:323-359 -> torch.distributions.MultivariateNormal) -----------------------------
    def mvn_parts(self, dist: Dist, scope):
        """-> (loc Expr, L leaf, W leaf, c leaf, d): the matrix argument is factorised once per matrix by MvnPrepOp; the
        density and the draw are then ordinary expressions over the cells."""
        args = {k: self.resolve_arg(dist.family, k, v, scope) for k, v in dist.args.items()}
        if dist.family == 'LowRankMultivariateNormal':
            return self._low_rank_parts(args)
        which = [k for k in MvnPrepOp.MODES if k in args]
        if len(which) != 1:
            raise Exception("Exactly one of covariance_matrix or precision_matrix or scale_tril may be specified.")
        S = self.materialize(self._prepare(args[which[0]]), tag=f'mvn:{which[0]}')
        if S.op != 'leaf' or len(S.pos_shape) != 2 or S.pos_shape[0] != S.pos_shape[1]:
            raise Exception(f"MultivariateNormal: {which[0]} must be a square matrix over the event dim "
                            f"(unnamed batch dims are not supported), got positional shape {S.pos_shape}")
        d = S.pos_shape[0]
        if d > 64:
            raise Exception("MultivariateNormal: event sizes above 64 are not supported by the device factorisation")
        if S.ref.id in self.needs:
            raise Exception(f"the gradient with respect to the {which[0]} of a MultivariateNormal is not provided "
                            f"(gradients flow to its value and loc)")
        if S.rename or S.mode:
            raise Exception("internal: the matrix argument of a MultivariateNormal must be a plain tensor")
        cache = getattr(self, '_mvn_cache', None)
        if cache is None:
            cache = self._mvn_cache = {}
        key = (S.ref.id, which[0])
        if key not in cache:
            axes = tuple(S.ref.axes)
            L = self.ws(axes, (d, d), name='mvn:L')
            W = self.ws(axes, (d, d), name='mvn:W')
            c = self.ws(axes, (), name='mvn:c')
            op = MvnPrepOp(S.ref, L, W, c, _prod(self.sizes[a] for a in axes), d, MvnPrepOp.MODES[which[0]])
            op.autodiff_as = 'skip'
            self.emit(op)
            cache[key] = (L, W, c)
        L, W, c = cache[key]
        lf = lambda pt: Expr.leaf(pt, pt.axes, pt.pos_shape)
        return self._prepare(args['loc']), lf(L), lf(W), lf(c), d

    def _low_rank_parts(self, args):
        """LowRankMultivariateNormal(loc, cov_factor [d, r], cov_diag [d]) (torch lowrank_multivariate_normal.py): the
        covariance cov_factor cov_factor^T + diag(cov_diag) is formed and factorised per matrix by MvnPrepOp (mode 3);
        torch evaluates the same density through the capacitance matrix (Woodbury), equal up to rounding."""
        Wf = self.materialize(self._prepare(args['cov_factor']), tag='mvn:cov_factor')
        Dg = self.materialize(self._prepare(args['cov_diag']), tag='mvn:cov_diag')
        if Wf.op != 'leaf' or Dg.op != 'leaf' or len(Wf.pos_shape) != 2 or Dg.pos_shape != Wf.pos_shape[:1]:
            raise Exception(f"LowRankMultivariateNormal: cov_factor must be a [d, r] matrix and cov_diag a [d] vector over "
                            f"the event dim (got positional shapes {Wf.pos_shape} and {Dg.pos_shape})")
        if tuple(Wf.ref.axes) != tuple(Dg.ref.axes):
            raise Exception(f"LowRankMultivariateNormal: cov_factor and cov_diag must vary over the same plates / K axes "
                            f"(got {Wf.ref.axes} and {Dg.ref.axes})")
        d, r = Wf.pos_shape
        if d > 64:
            raise Exception("LowRankMultivariateNormal: event sizes above 64 are not supported by the device factorisation")
        if Wf.ref.id in self.needs or Dg.ref.id in self.needs:
            raise Exception("the gradient with respect to cov_factor / cov_diag of a LowRankMultivariateNormal is not "
                            "provided (gradients flow to its value and loc)")
        if Wf.rename or Wf.mode or Dg.rename or Dg.mode:
            raise Exception("internal: the matrix arguments of a LowRankMultivariateNormal must be plain tensors")
        cache = getattr(self, '_mvn_cache', None)
        if cache is None:
            cache = self._mvn_cache = {}
        key = (Wf.ref.id, Dg.ref.id, 'low_rank')
        if key not in cache:
            axes = tuple(Wf.ref.axes)
            L = self.ws(axes, (d, d), name='mvn:L')
            W = self.ws(axes, (d, d), name='mvn:W')
            c = self.ws(axes, (), name='mvn:c')
            op = MvnPrepOp(Wf.ref, L, W, c, _prod(self.sizes[a] for a in axes), d, 3, low_rank=(Dg.ref, r))
            op.autodiff_as = 'skip'
            self.emit(op)
            cache[key] = (L, W, c)
        L, W, c = cache[key]
        lf = lambda pt: Expr.leaf(pt, pt.axes, pt.pos_shape)
        return self._prepare(args['loc']), lf(L), lf(W), lf(c), d

    def _mvn_density(self, dist, value, scope, tag) -> PT:
        loc, L, W, c, d = self.mvn_parts(dist, scope)
        value = self._prepare(value)
        if value.pos_shape[-1:] != (d,) or len(value.pos_shape) != 1 or loc.pos_shape not in ((d,), ()):
            raise Exception(f"MultivariateNormal: value / loc must be vectors of the event size {d} "
                            f"(got {value.pos_shape} and {loc.pos_shape})")
        mk = Expr.make
        z = mk('sumlast', mk('mul', W, mk('sub', value, loc)))            # W (x - loc): [cells..., d]
        body = mk('sub', c, mk('mul', Expr.const(0.5), mk('sumlast', mk('square', z))))
        return self.emit_expr(self._prepare(body), nred='all', tag=tag)

    def _composed_density(self, dist, value, scope, tag) -> PT:
        """Families of COMPOSED: the torch log_prob restated over traced operands, lowered like any model lambda."""
        order, fn = COMPOSED[dist.family]
        args = {k: Proxy(self.resolve_arg(dist.family, k, v, scope)) for k, v in dist.args.items()}
        vec = dist.family in _VECTOR_FAMILIES
        both = 'logits' in order and 'probs' in order                    # ContinuousBernoulli uses the two of them
        if 'logits' in order and 'logits' not in args:
            p = args['probs'] if both else args.pop('probs')
            # binary families: logit(p); vector families: log of the normalised probabilities (categorical.py:60-65)
            args['logits'] = (p.log() - p.sum(-1).log()) if vec else _probs_to_logits(p)
        if 'probs' in order and 'probs' not in args:
            lg = args['logits'] if both else args.pop('logits')
            args['probs'] = lg.sigmoid()                                  # logits_to_probs(is_binary=True)
        v = Proxy(value)
        extra = []
        if dist.family == 'Categorical':
            # value = the class index (a scalar per cell), probs / logits = one vector over the last positional dim
            lshape = args['logits'].expr.pos_shape
            if value.pos_shape != () or len(lshape) != 1:
                raise Exception(f"Categorical: the value must be a scalar class index and probs / logits one vector "
                                f"(got positional shapes {value.pos_shape} and {lshape})")
            iota = self.const_input(torch.arange(lshape[0], dtype=torch.float64))
            extra = [Proxy(Expr.leaf(iota, (), iota.pos_shape))]
        elif vec and (len(value.pos_shape) < 1 or args['logits'].expr.pos_shape[-1:] != value.pos_shape[-1:]):
            raise Exception(f"{dist.family}: value and probs / logits must be vectors of one size over the last "
                            f"positional dim (got {value.pos_shape} and {args['logits'].expr.pos_shape})")
        if dist.family in ('ContinuousBernoulli', 'VonMises'):
            # the normaliser depends on the parameter only and is longer than one VM program: its pieces are
            # materialised (hoisted) tensors of the parameter's shape, like any `arg` of a density
            extra = [lambda e: Proxy(self.materialize(self._prepare(e.expr), tag='arg'))]
        body = fn(v, *[args[k] for k in order], *extra).expr
        return self.emit_expr(self._prepare(body), nred='all', tag=tag)

    def _dirichlet_density(self, dist, value, scope, tag) -> PT:
        """torch.distributions.Dirichlet.log_prob over the last positional dim (the simplex), as expressions:
        sum_j xlogy(alpha_j - 1, x_j) + lgamma(sum_j alpha_j) - sum_j lgamma(alpha_j)."""
        alpha = self._prepare(self.resolve_arg(dist.family, 'concentration', dist.args['concentration'], scope))
        value = self._prepare(value)
        if len(value.pos_shape) != 1 or len(alpha.pos_shape) != 1 or alpha.pos_shape != value.pos_shape:
            raise Exception(f"Dirichlet: value and concentration must be vectors of one size "
                            f"(got {value.pos_shape} and {alpha.pos_shape})")
        mk = Expr.make
        body = mk('sub', mk('add', mk('sumlast', mk('mul', mk('sub', alpha, Expr.const(1.0)), mk('log', value))),
                            mk('lgamma', mk('sumlast', alpha))),
                  mk('sumlast', mk('lgamma', alpha)))
        return self.emit_expr(self._prepare(body), nred='all', tag=tag)

    FAN_EVENT_EXTENTS = (1, 2, 3, 4, 6, 8, 12, 16, 18, 24, 32)

    def _try_normal_fan(self, opname, operands, body, out, tag):
        """Pattern of csrc/fused.cuh normal_fan_kernel: Normal whose value/loc carry the row axes and
        whose scale carries (at most) one K axis of its own."""
        if opname != 'Normal' or len(body.pos_shape) > 1:
            return None
        D = body.pos_shape[0] if body.pos_shape else 1
        if D not in self.FAN_EVENT_EXTENTS:
            return None
        leaves = []
        for e in operands:
            if e.op == 'const':
                e = Expr.leaf(self.const_input(torch.tensor(e.value, dtype=torch.float64)), (), ())
            if e.op != 'leaf' or e.mode != 0 or e.rename or e.pos_shape not in ((), (D,)):
                return None
            leaves.append(e)
        v, l, sc = leaves
        fan_axes = [a for a in sc.axes if a not in v.axes and a not in l.axes]
        if len(fan_axes) > 1 or any(a not in fan_axes for a in sc.axes):
            return None
        row_axes = [a for a in out.axes if a not in fan_axes]
        n_rows = _prod(self.sizes[a] for a in row_axes)
        F = self.sizes[fan_axes[0]] if fan_axes else 1
        if n_rows * F < 4096 or len(row_axes) > MAXD:
            return None
        return NormalFanOp(out, D, [self.axdim(a) for a in row_axes], plain(v.ref), plain(l.ref), plain(sc.ref),
                           fan_axes[0] if fan_axes else None, F, tag)

    # -- plate recursion (logpq.py:68-155, 257-332) -------------------------------------------
    def plan_plate(self, name, P: Plate, Q: Plate, active, scope):
        if name is not None:
            active = (*active, name)
        scope = dict(scope)
        lfs = []
        for key, e in self.extra_factors:
            plates = set(a for a in e.axes if a in self.all_plates)
            if plates == set(active):
                lfs.append(self._extra_factor(key, e))
        Knon, Kts, Kinits = [], [], []
        for childname, childQ in Q.grouped_prog.items():
            if isinstance(childQ, dict):
                childP = {v: P.flat_prog[v] for v in childQ}
                lf, a, b, c = self.plan_group(childname, childP, childQ, active, scope)
            else:
                lf = self.plan_plate(childname, P.flat_prog[childname], childQ, active, scope)
                a = b = c = ()
            lfs.append(lf)
            Knon.extend(a); Kts.extend(b); Kinits.extend(c)
        level_steps = []
        self.level_factors[tuple(active)] = (list(lfs), tuple(Knon))
        lf = self.contract(lfs, tuple(Knon), active, level_steps)
        self.level_steps[tuple(active)] = (level_steps, Q)
        if name is None:
            return lf
        if Kinits:
            return self.chain(lf, name, Kinits[0], Kts[0])
        return self.plate_sum(lf, name)

    def _extra_factor(self, key, e: Expr) -> LogicalFactor:
        if len(e.pos_shape) > 0 or e.op != 'leaf':
            pt = self.emit_expr(self._prepare(e), nred='all', tag=f'elf:{key}')
        else:
            pt = e.ref
        return LogicalFactor([(plain(pt), 1.0)], 0.0, tuple(pt.axes))

    def plan_group(self, name, prog_P, prog_Q, active, scope):
        if datagroup(prog_Q):
            k = next(iter(prog_Q))
            if k not in self.sig or self.sig[k].role != 'data':
                raise Exception(f"no data tensor was provided for {k}")
            s = self.sig[k]
            value = Expr.leaf(self.inputs[k], s.axes, s.pos_shape)
            pt = self.density(prog_P[k], value, scope, tag=f'logP:{k}')
            return LogicalFactor([(plain(pt), 1.0)], 0.0, pt.axes), (), (), ()

        K_axis = Kname(name)
        K = self.sizes[K_axis]
        T_axis = active[-1] if active else None
        tensors, q_tensors, Kinits = [], [], []
        for k in prog_P:
            dP, dQ = prog_P[k], prog_Q[k]
            if k not in self.sig or self.sig[k].role != 'sample':
                raise Exception(f"no sample was provided for latent variable {k}")
            s = self.sig[k]
            value = Expr.leaf(self.inputs[k], s.axes, s.pos_shape)
            for which, d, store in (('P', dP, tensors), ('Q', dQ, q_tensors)):
                sc = scope
                if isinstance(d, Timeseries):
                    sc, Kinit = self._timeseries_scope(d, k, value, scope, T_axis, K_axis)
                    if which == 'P':
                        Kinits.append(Kinit)
                    d = d.trans
                pt = self.density(d, value, sc, tag=f'log{which}:{k}')
                store.append(pt)
            scope[k] = value                     # later members of the group may refer to it
        # mixture-Q reduction over parent Ks (Sampler.py:118-134)
        q_axes = _union_axes([pt.axes for pt in q_tensors])
        parents = tuple(a for a in q_axes if a != K_axis and a not in active)
        facs = [(plain(pt), 1.0) for pt in tensors]
        if parents:
            rd = [self.axdim(a) for a in parents]
            out_axes = self.canon_order([a for a in q_axes if a not in parents])
            out = self.ws(out_axes, name=f'logQ~:{name}')
            cadd = -sum(math.log(self.sizes[a]) for a in parents)
            self.emit(ReduceOp(R_LSE_EPS, out, [self.axdim(a) for a in out_axes], rd,
                                     [(plain(pt), 1.0) for pt in q_tensors], cadd=cadd, tag=f'reduce_logQ:{name}'))
            facs.append((plain(out), -1.0))
        else:
            facs.extend((plain(pt), -1.0) for pt in q_tensors)
        axes = self.canon_order(_union_axes([lf.pt.axes for lf, _ in facs]))
        lf = LogicalFactor(facs, -math.log(K), axes)
        if Kinits:
            return lf, (), (K_axis,), (Kinits[0],)
        return lf, (K_axis,), (), ()

    def _timeseries_scope(self, ts: Timeseries, varname, value: Expr, scope, T_axis, K_axis):
        """Timeseries.py:203-245: prev[t] = x[t-1] with K renamed to the init variable's K; prev[0] = init."""
        if ts.init not in scope:
            raise Exception(f"Timeseries initial state {ts.init} is not in scope")
        init = scope[ts.init]
        Kinit = Kname(self.v2g[ts.init])
        if Kinit not in init.axes or T_axis in init.axes:
            raise Exception("Timeseries initial state must be a latent of the immediately enclosing plate")
        prev_axes = tuple(Kinit if a == K_axis else a for a in value.axes)
        shifted = Expr.leaf(value.ref, prev_axes, value.pos_shape, rename={Kinit: K_axis}, mode=1, mdim=T_axis)
        first = Expr.leaf(init.ref, init.axes, init.pos_shape, mode=2, mdim=T_axis)
        first.axes = tuple(init.axes) + (T_axis,)         # the masked load depends on t
        prev = Expr.make('add', shifted, first)
        return {**scope, 'prev': prev}, Kinit

    # -- global importance sampling baseline (SampleNonMP.py:127-203 `non_mp_log_prob`) ---------------
    def plan_nonmp(self, name, P: Plate, Q: Plate, active, scope):
        """One plate level of the non-massively-parallel log-probability: every sample carries the single axis
        NONMP_K, every factor is `[active plates..., K]`, nothing is contracted: log P - log Q per variable, summed
        over the plates level by level (the plate sums take the same fused paths as the MP plan)."""
        if name is not None:
            active = (*active, name)
        scope = dict(scope)
        facs, const = [], 0.0
        for key, e in self.extra_factors:
            plates = set(a for a in e.axes if a in self.all_plates)
            if plates == set(active):
                lf = self._extra_factor(key, e)
                facs.extend(lf.tensors)
        for k, dQ in Q.flat_prog.items():
            dP = P.flat_prog[k]
            if isinstance(dP, Timeseries) or isinstance(dQ, Timeseries):
                raise Exception("SampleNonMP does not support a Timeseries (reference SampleNonMP.py:156)")
            if isinstance(dQ, Plate):
                lf = self.plan_nonmp(k, dP, dQ, active, scope)
                facs.extend(lf.tensors); const += lf.const
                continue
            if isinstance(dQ, Data):
                if k not in self.sig or self.sig[k].role != 'data':
                    raise Exception(f"no data tensor was provided for {k}")
                s = self.sig[k]
                value = Expr.leaf(self.inputs[k], s.axes, s.pos_shape)
                pts = [(self.density(dP, value, scope, tag=f'logP:{k}'), 1.0)]
            else:
                if k not in self.sig or self.sig[k].role != 'sample':
                    raise Exception(f"no sample was provided for latent variable {k}")
                s = self.sig[k]
                value = Expr.leaf(self.inputs[k], s.axes, s.pos_shape)
                pts = [(self.density(dP, value, scope, tag=f'logP:{k}'), 1.0),
                       (self.density(dQ, value, scope, tag=f'logQ:{k}'), -1.0)]
                scope[k] = value
            for pt, coeff in pts:
                missing = [a for a in active if a not in pt.axes]
                if missing:
                    # the reference asserts dims == active plates + K (SampleNonMP.py:176,186-187)
                    raise Exception(f"non-MP factor of {k} does not span the plates {missing}")
                facs.append((plain(pt), coeff))
        if not facs:
            raise Exception("plate without factors")
        axes = self.canon_order(_union_axes([lf.pt.axes for lf, _ in facs]))
        lf = LogicalFactor(facs, const, axes)
        if name is None:
            return lf
        return self.plate_sum(lf, name)

    def build_sampling_nonmp(self):
        """SampleNonMP._importance_sample_idxs (SampleNonMP.py:71-90): N draws from the categorical over the one K axis
        with weights exp(lpq - max): one SampleOp on the [K] vector the forward pass left in the workspace."""
        if self.N is None:
            raise Exception("resampling program needs the number of posterior samples N")
        plan = self.plan
        plan.N = self.N
        sizes = dict(self.sizes)
        sizes['N'] = self.N
        idx = PT(('N',), (), sizes, 'output', index=0, name='idx')
        plan.sample_groups = [(NONMP_K, ())]
        u = PT(('N',), (), sizes, 'aux', index=0, name='u0')
        plan.sample_steps = [((), (NONMP_K,))]
        nd = [('ax', 'N', self.N)]
        return [SampleOp(nd, [self.axdim(NONMP_K)], [(lf, coeff, []) for lf, coeff in self.nonmp_lf.tensors], [],
                         (u, nd), [idx])]

    # -- contraction (reduce_Ks.py:236-298) --------------------------------------------------
    def contract(self, lfs, Ks_to_sum, active, level_steps):
        if not lfs:
            raise Exception("plate without factors")
        path = greedy_path([lf.axes for lf in lfs], Ks_to_sum, self.sizes)
        self.level_paths[tuple(active)] = path
        lfs = list(lfs)
        for idxs in path:
            chosen = [lfs[i] for i in idxs]
            lfs = [lfs[i] for i in range(len(lfs)) if i not in idxs]
            remaining = set(a for lf in lfs for a in lf.axes)
            chosen_axes = _union_axes([lf.axes for lf in chosen])
            ks = tuple(k for k in Ks_to_sum if k in chosen_axes and k not in remaining)
            tensors = [tc for lf in chosen for tc in lf.tensors]
            const = sum(lf.const for lf in chosen)
            if not ks:
                lfs.append(LogicalFactor(tensors, const, self.canon_order(chosen_axes)))
                continue
            out_axes = self.canon_order([a for a in chosen_axes if a not in ks])
            out = self.ws(out_axes, name='lse[' + ','.join(ks) + ']')
            red = ReduceOp(R_LSE_EPS, out, [self.axdim(a) for a in out_axes],
                           [self.axdim(a) for a in ks], tensors, cadd=const, tag='contract:' + ','.join(ks))
            fused = self._try_fan_lse(red, tensors, ks, const) if self.fast_paths else None
            self.emit(fused if fused is not None else red)
            level_steps.append(Step(tuple(active), tensors, ks))
            lfs.append(LogicalFactor([(plain(out), 1.0)], 0.0, out_axes))
        assert len(lfs) == 1
        return lfs[0]

    def _try_fan_lse(self, red, tensors, ks, const):
        """Fuse a normal_fan factor into the LSE step that consumes it (csrc/fused.cuh fan_lse_kernel)
        when nobody else needs the materialised factor: no resampling program, no adjoint of the
        factor itself (RWS / marginals of other groups only need the small factors' adjoints)."""
        if len(ks) != 1 or getattr(self, 'with_sample', False):
            return None
        kappa = ks[0]
        cands = [(lf, c) for lf, c in tensors if lf.pt.id in self.fan_by_out and c == 1.0 and type(lf) is LeafRef]
        if len(cands) != 1:
            return None
        big, _ = cands[0]
        fan = self.fan_by_out[big.pt.id]
        if fan not in self.fwd or big.pt.id in self.needs or fan.fan_axis is None or fan.F < 8:
            return None
        row_axes = [d[1] for d in fan.rows]
        if kappa not in row_axes or self.sizes[kappa] > 128 or kappa == fan.fan_axis:
            return None
        small = [(lf, c) for lf, c in tensors if lf is not big]
        for lf, c in small:
            if type(lf) is not LeafRef or lf.rename or lf.mode or lf.pt.pos_shape:
                return None
            if any(a not in row_axes for a in lf.pt.axes):
                return None
        if fan.v.stride(self.axdim(kappa)) == 0 and fan.l.stride(self.axdim(kappa)) == 0:
            return None
        rho = [d for d in fan.rows if d[1] != kappa]
        # shared-memory footprint of csrc/fused.cuh fan_lse2_kernel with one rho per warp (its minimum)
        DP, KP = (fan.D + 5) // 4 * 4, self.sizes[kappa] + 4
        tile = KP * DP + 32
        smem = (4 * (tile + 32 * (KP + 1)) + 32 * 6 * DP) * self.itemsize + 4 * 13 * 8
        if smem > 190 * 1024 or tile * self.itemsize > 44 * 1024:
            return None
        self.fwd.remove(fan)
        op = FanLseOp(red.out, fan.D, rho, self.axdim(kappa), fan.v, fan.l, fan.s, fan.fan_axis, fan.F,
                      small, const, fan.autodiff_as, red, tag='fan_lse:' + fan.tag)
        # commit the adjoint to the dense tensor-core kernel (compact gS: one slot per fan group instead of one per
        # lam) unless the caller asked for the other kernels when the plan is built
        if not (os.environ.get("ALAN_B200_NO_TC") or os.environ.get("ALAN_B200_TC_BLOCKDIAG")):
            op.dense = dense_fan_geometry(op, self.itemsize)
        # Opt-in (ALAN_B200_QFUSE=1 when the plan is built): measured on B200 at cfg-5 the builder warps of the dense
        # kernel are on its critical path (the MMA issuer waits on them ~25 % of the time), so the ~350 cycles per block
        # the inline Q factor adds to them cost as much (+9 us forward, +12 us adjoint) as the removed pass (23 us).
        if op.dense is not None and os.environ.get("ALAN_B200_QFUSE") == "1":
            self._try_inline_q(op)
        return op

    def _normal3_parts(self, E):
        """(value, loc, scale) leaves of an ExprOp that is exactly `sum_d log N(value; loc, scale)` over plain leaves."""
        if not isinstance(E, ExprOp) or E.acc or E.scale != 1.0 or len(E.red) > 1 or any(d[0] != 'ax' for d in E.keep):
            return None
        code = E.codeobj
        if [ins[0] for ins in code.instrs] != [VOPS['load']] * 3 + [VOPS['Normal']] or len(code.leaves) != 3:
            return None
        _, dst, ra, rb, rc, _ = code.instrs[3]
        if code.res != dst:
            return None
        leaf_of = {code.instrs[i][1]: code.leaves[code.instrs[i][2]] for i in range(3)}
        parts = (leaf_of[ra], leaf_of[rb], leaf_of[rc])
        if any(type(x) is not LeafRef or x.rename or x.mode for x in parts):
            return None
        return parts

    def _try_inline_q(self, op: 'FanLseOp'):
        """The dense tensor-core kernel's builder warps hold every value row v[u, kappa, :] in registers.  If one of
        the small factors of the contraction is the Gaussian Q factor of the SAME value tensor with per-user loc and
        scale (`sum_d log N(v; loc[u,d], scale[u,d])`, nobody else reading it), it is evaluated there: the factor's own
        pass over v (and its [u, kappa] tensor) disappears from the forward program."""
        lam = op.dense[0]
        ev = ('ev', 0, op.D)
        for i, (lf, c) in enumerate(op.bfactors):
            E = self.producer.get(lf.pt.id)
            parts = self._normal3_parts(E)
            if parts is None or E not in self.fwd or lf.pt.id in self.needs_materialised:
                continue
            v, l, sc = parts
            if v.pt is not op.v.pt or (E.red[0][2] if E.red else 1) != op.D or op.D > 32:
                continue
            if [v.stride(d) for d in op.rho + [op.kappa, ev]] != [op.v.stride(d) for d in op.rho + [op.kappa, ev]]:
                continue
            if any(x.stride(op.kappa) != 0 or x.stride(lam) != 0 or (op.D > 1 and x.stride(ev) == 0) for x in (l, sc)):
                continue
            if set((d[0], d[1]) for d in E.keep) != set((d[0], d[1]) for d in op.rho + [op.kappa] if d != lam):
                continue
            self.fwd.remove(E)
            op.qterm = (E, l, sc, c)
            op.bfactors = op.bfactors[:i] + op.bfactors[i + 1:]
            return

    def plate_sum(self, lf: LogicalFactor, plate):
        out_axes = tuple(a for a in lf.axes if a != plate)
        out = self.ws(out_axes, name=f'sum[{plate}]')
        n = self.sizes[plate]
        od = [self.axdim(a) for a in out_axes]
        rd = [self.axdim(plate)]
        n_out = max(1, _prod(d[2] for d in od))
        nsplit = _choose_split(n_out, n)
        fused = self._try_bern_dot_sum(lf, out, od, rd, n) if (self.fast_paths and nsplit == 1) else None
        if fused is None and self.fast_paths:
            fused = self._try_normal_poly_sum(lf, out, od, rd, n)
        fan = self._fan_for_plate_sum(lf, od, rd) if (self.fast_paths and nsplit > 1) else None
        if fused is not None:
            self.emit(fused)
        elif fan is not None:
            # the dense fan_lse kernel already holds every out[user, lam, f] in a register: it keeps per-(CTA, team)
            # running sums over its users and writes them as one partial row per CTA; only those are summed here
            part = self.ws_raw(FAN_PSUM_ROWS * n_out, name=f'partial[{plate}]')
            fan.psum = (part, FAN_PSUM_ROWS, od)
            second = ReduceOp(R_SUM, out, od, [('sp', 0, FAN_PSUM_ROWS)], [(_PartialRef(part, od, FAN_PSUM_ROWS), 1.0)],
                              cadd=lf.const * n, tag=f'plate_sum:{plate}')
            second.autodiff_as = ReduceOp(R_SUM, out, od, rd, lf.tensors, cadd=lf.const * n)
            self.emit(second)
        elif nsplit > 1:
            part = self.ws_raw(nsplit * n_out, name=f'partial[{plate}]')
            first = ReduceOp(R_SUM, part, od, rd, lf.tensors, nsplit=nsplit, tag=f'plate_sum_partial:{plate}')
            first.autodiff_as = 'skip'
            self.emit(first)
            sd = ('sp', 0, nsplit)
            second = ReduceOp(R_SUM, out, od, [sd], [(_PartialRef(part, od, nsplit), 1.0)],
                              cadd=lf.const * n, tag=f'plate_sum:{plate}')
            # the adjoint is derived from the unsplit form of the same sum
            second.autodiff_as = ReduceOp(R_SUM, out, od, rd, lf.tensors, cadd=lf.const * n)
            self.emit(second)
        else:
            self.emit(ReduceOp(R_SUM, out, od, rd, lf.tensors, cadd=lf.const * n, tag=f'plate_sum:{plate}'))
        if plate == self.shard_plate:
            self.plan.allreduce = out
            if self.fused_collectives:
                self.fwd.append(XReduceOp(0, [out]))
            else:
                self.fwd_segments.append(self.fwd)
                self.fwd = []
        return LogicalFactor([(plain(out), 1.0)], 0.0, out_axes)

    def _fan_for_plate_sum(self, lf, od, rd):
        """The FanLseOp whose output is the only term of this plate sum, when it runs on the dense tensor-core
        kernel and the summed plate is exactly its user axis (so the sum can ride in that kernel's epilogue)."""
        if len(lf.tensors) != 1:
            return None
        ref, coeff = lf.tensors[0]
        if coeff != 1.0 or type(ref) is not LeafRef or ref.rename or ref.mode:
            return None
        op = self.producer.get(ref.pt.id)
        if not isinstance(op, FanLseOp) or op.dense is None or op.psum is not None or op not in self.fwd:
            return None
        lam = op.dense[0]
        users = [d for d in op.rho if d != lam]
        fdim = ('ax', op.fan_axis, op.F)
        if [(d[0], d[1]) for d in users] != [(d[0], d[1]) for d in rd]:
            return None
        if sorted((d[0], d[1]) for d in od) != sorted((d[0], d[1]) for d in (lam, fdim)):
            return None
        return op

    def _try_normal_poly_sum(self, lf, out, od, rd, n):
        """Fuse  Normal(value; polynomial loc, scale)  ->  plate sum  into csrc/normal_poly.cuh when the factor is read
        by nobody else and no gradient flows through it (marginals, moments, RWS, resampling; the reparameterised
        path keeps the generic expression and its adjoint).  The polynomial is recovered from the traced program."""
        if len(lf.tensors) != 1 or len(rd) != 1:
            return None
        ref, coeff = lf.tensors[0]
        if coeff != 1.0 or type(ref) is not LeafRef or ref.rename or ref.mode:
            return None
        E = self.producer.get(ref.pt.id)
        if not isinstance(E, ExprOp) or E not in self.fwd or E.red or E.acc or E.scale != 1.0:
            return None
        if ref.pt.id in self.needs or ref.pt.id in self.needs_materialised:
            return None
        if set((d[0], d[1]) for d in E.keep) != set((d[0], d[1]) for d in od + rd) or any(d[0] != 'ax' for d in E.keep):
            return None
        code = E.codeobj
        if any(type(x) is not LeafRef or x.rename or x.mode for x in code.leaves):
            return None
        if any(x.pt.id in self.needs for x in code.leaves):
            return None
        op0, dst, ra, rb, rc, _ = code.instrs[-1]
        if op0 != VOPS['Normal'] or code.res != dst:
            return None
        # registers as polynomials over the leaves: {sorted tuple of leaf indices: coefficient}
        poly = {}
        for (op, d, a, b, c, _) in code.instrs[:-1]:
            if op == VOPS['load']:
                poly[d] = {(a,): 1.0}
            elif op == VOPS['const']:
                poly[d] = {(): float(code.consts[a])}
            elif op in (VOPS['add'], VOPS['sub']):
                if a not in poly or b not in poly:
                    return None
                out_p = dict(poly[a])
                sgn = 1.0 if op == VOPS['add'] else -1.0
                for m, cf in poly[b].items():
                    out_p[m] = out_p.get(m, 0.0) + sgn * cf
                poly[d] = out_p
            elif op == VOPS['mul']:
                if a not in poly or b not in poly:
                    return None
                out_p = {}
                for m1, c1 in poly[a].items():
                    for m2, c2 in poly[b].items():
                        m = tuple(sorted(m1 + m2))
                        out_p[m] = out_p.get(m, 0.0) + c1 * c2
                poly[d] = out_p
            elif op == VOPS['neg']:
                if a not in poly:
                    return None
                poly[d] = {m: -cf for m, cf in poly[a].items()}
            elif op == VOPS['mov']:
                if a not in poly:
                    return None
                poly[d] = dict(poly[a])
            else:
                return None
        if ra not in poly or rb not in poly or rc not in poly:
            return None
        resid = dict(poly[ra])
        for m, cf in poly[rb].items():
            resid[m] = resid.get(m, 0.0) - cf
        resid = {m: cf for m, cf in resid.items() if cf != 0.0}
        sc = poly[rc]
        zdim = rd[0]
        is_z = lambda li: code.leaves[li].stride(zdim) != 0
        # the scale: one K leaf or a constant
        if len(sc) != 1:
            return None
        (sm, scf), = sc.items()
        if len(sm) > 1 or (len(sm) == 1 and (scf != 1.0 or is_z(sm[0]))):
            return None
        zl, kl = [], []                        # leaf indices of the code, in kernel order

        def slot(lst, li):
            if li not in lst:
                lst.append(li)
            return lst.index(li)
        zterms, kterms = [], []
        for m, cf in resid.items():
            zs = [li for li in m if is_z(li)]
            ks = [li for li in m if not is_z(li)]
            if len(zs) > 2 or len(ks) > 2:
                return None
            if zs:
                zterms.append((cf, [slot(zl, li) for li in zs], [slot(kl, li) for li in ks]))
            else:
                kterms.append((cf, [], [slot(kl, li) for li in ks]))
        scale_leaf, scale_const = (-1, float(scf)) if len(sm) == 0 else (slot(kl, sm[0]), 0.0)
        if not (1 <= len(zterms) <= 6) or len(kterms) > 8 or len(zl) > 8 or len(kl) > 8:
            return None
        # rows = the output dims some Z leaf walks; K dims = the others (no Z leaf may depend on them, by construction)
        zleaves, kleaves = [code.leaves[li] for li in zl], [code.leaves[li] for li in kl]
        rows = [d for d in od if any(x.stride(d) != 0 for x in zleaves)]
        kd = [d for d in od if d not in rows]
        if len(rows) + len(kd) + 1 > MAXD:
            return None
        R = ReduceOp(R_SUM, out, od, rd, lf.tensors, cadd=lf.const * n)
        self.fwd.remove(E)
        return NormalPolySumOp(out, rows, kd, [zdim], zleaves, kleaves, zterms, kterms, scale_leaf, scale_const,
                               lf.const * n, [E, R], tag='normal_poly_sum:' + E.tag)

    def _try_bern_dot_sum(self, lf, out, od, rd, n):
        """Fuse  dot -> Bernoulli(logits) -> plate sum  into csrc/fused.cuh bern_dot_sum_kernel when nobody
        else needs the intermediates (no resampling program, no gradient through the factor)."""
        if getattr(self, 'with_sample', False) or len(lf.tensors) != 1:
            return None
        ref, coeff = lf.tensors[0]
        if coeff != 1.0 or type(ref) is not LeafRef or ref.rename or ref.mode:
            return None
        E = self.producer.get(ref.pt.id)
        if not isinstance(E, ExprOp) or E not in self.fwd or E.red or E.acc or E.scale != 1.0:
            return None
        code = E.codeobj
        ops = [ins[0] for ins in code.instrs]
        if ops != [VOPS['load'], VOPS['load'], VOPS['Bernoulli_logits']] or len(code.leaves) != 2:
            return None
        _, _, ra, rb, _, _ = code.instrs[2]
        leaf_of = {code.instrs[i][1]: code.leaves[code.instrs[i][2]] for i in (0, 1)}
        y, lg = leaf_of[ra], leaf_of[rb]
        Dt = self.producer.get(lg.pt.id)
        if not isinstance(Dt, DotOp) or Dt not in self.fwd or len(Dt.red) != 1 or lg.rename or lg.mode or y.rename or y.mode:
            return None
        D = Dt.red[0][2]
        if D not in self.FAN_EVENT_EXTENTS:
            return None
        if any(x.id in self.needs for x in (ref.pt, lg.pt, Dt.a.pt, Dt.b.pt, y.pt)):
            return None
        a, b = Dt.a, Dt.b
        if any(a.stride(d) != 0 for d in rd):
            a, b = b, a
        if any(a.stride(d) != 0 for d in rd):
            return None
        if set((d[0], d[1]) for d in E.keep) != set((d[0], d[1]) for d in od + rd):
            return None
        R = ReduceOp(R_SUM, out, od, rd, lf.tensors, cadd=lf.const * n)
        self.fwd.remove(E)
        self.fwd.remove(Dt)
        return BernDotSumOp(out, od, rd, D, a, b, y, lf.const * n, [Dt, E, R], tag='bern_dot_sum:' + E.tag)

    def fuse_side_factors(self):
        """Run after the adjoint programs have been derived (they see the unfused ops).  The Bernoulli-dot-sum kernel
        keeps every row a[o, :] (the sample z[u, k, :] at cfg-2 / cfg-5) in registers.  A Gaussian factor of the same
        rows with per-row-group loc and scale, `E.out[o] = sum_e log N(a[o, e]; loc, scale)` -- the mean-field Q factor
        logQ(z) -- becomes a second output of that kernel: its own pass over `a` disappears from the forward program
        (cfg-5: 23 us and 21.6 MB of reads).  E stays the op the adjoint was derived from."""
        for seg in self.plan.programs[:self.plan.n_fwd]:
            for B in [o for o in seg if isinstance(o, BernDotSumOp) and o.side is None]:
                ev = ('ev', 0, B.D)
                for E in [o for o in seg if isinstance(o, ExprOp)]:
                    parts = self._normal3_parts(E)
                    if parts is None or (E.red[0][2] if E.red else 1) != B.D or E.out.space != 'ws':
                        continue
                    v, l, sc = parts
                    keyset = lambda ds: set((d[0], d[1]) for d in ds)
                    if v.pt is not B.a.pt or keyset(E.keep) != keyset(B.od):
                        continue
                    if [v.stride(d) for d in B.od + [ev]] != [B.a.stride(d) for d in B.od + [ev]]:
                        continue
                    if [plain(E.out).stride(d) for d in B.od] != [plain(B.out).stride(d) for d in B.od]:
                        continue
                    if B.D > 1 and (l.stride(ev) == 0 or sc.stride(ev) == 0):
                        continue
                    iB, iE = seg.index(B), seg.index(E)
                    lo, hi = min(iB, iE), max(iB, iE)
                    first = seg[lo]
                    if any(first.out.id in (x.id for x in self.op_inputs(o)) for o in seg[lo + 1:hi]):
                        continue                    # somebody between the two reads the earlier one's output
                    B.side = (E, l, sc)
                    seg[hi] = B
                    del seg[lo]
                    break

    def chain(self, lf: LogicalFactor, T_axis, Kinit, Kts):
        """logpq.py:131-143: order to [T, Kprev, Kcurr] (other axes batch), chain_logmmexp, logsumexp."""
        outer = tuple(a for a in lf.axes if a not in (T_axis, Kinit, Kts))
        ms_axes = outer + (T_axis, Kinit, Kts)
        for a in (T_axis, Kinit, Kts):
            if a not in lf.axes:
                raise Exception(f"Timeseries factor lacks axis {a}")
        if self.sizes[Kinit] != self.sizes[Kts]:
            raise Exception("Timeseries needs the same K for the initial state and the chain")
        ms = self.ws(ms_axes, name='chain_ms')
        od = [self.axdim(a) for a in ms_axes]
        self.emit(ReduceOp(R_SUM, ms, od, [], lf.tensors, cadd=lf.const, tag='chain_ms'))
        K, T = self.sizes[Kts], self.sizes[T_axis]
        n_outer = _prod(self.sizes[a] for a in outer)
        n, tot = T, 0
        while n > 1:
            n = n // 2 + n % 2
            tot += n_outer * n * K * K
        levels = self.ws_raw(max(tot, 1), name='chain_levels')
        out = self.ws(outer + (Kinit,), name='chain_out')
        self.emit(ChainOp(ms, levels, out, n_outer, T, K))
        return LogicalFactor([(plain(out), 1.0)], 0.0, out.axes)

    # -- top level ----------------------------------------------------------------------------
    def _last_contraction_as_root(self, lf):
        """The op that may write the log-evidence itself: `lp = t + const` with `t` the scalar output of the reduction
        emitted last (the top-level contraction over the last K axis), which nobody else reads."""
        if len(lf.tensors) != 1 or lf.tensors[0][1] != 1.0 or os.environ.get("ALAN_B200_NO_TAILFOLD") == "1":
            return None
        ref = lf.tensors[0][0]
        if type(ref) is not LeafRef or ref.rename or ref.mode or not self.fwd:
            return None
        op = self.fwd[-1]
        if not isinstance(op, ReduceOp) or op.out is not ref.pt or op.out.space != 'ws' or op.out.shape != () \
                or op.nsplit != 1 or op.acc or op.scale != 1.0 or op.mode not in (R_SUM, R_LSE_EPS, R_LSE) \
                or getattr(op, 'autodiff_as', None) is not None or ref.pt.id in self.needs_materialised:
            return None
        return op

    def build(self, grad_names=(), with_sample=False) -> Plan:
        if grad_names or not self.grad_names:
            self.set_grad_names(grad_names)
        self.with_sample = with_sample
        # The log-evidence is written straight into output 0 by the op that produces it: the microsecond-scale ops at
        # the end of the forward program are a chain of DEPENDENT launches (~4 us each under graph replay,
        # profiles/r02_small_ops.md), so neither a copy (`lp_out`) nor a one-term sum (`lp`) follows the last contraction.
        lp = PT((), (), self.sizes, 'output', index=0, name='lp')
        self.lp_ws = lp                          # the root of the adjoint program (build_backward seeds its adjoint)
        if self.nonmp:
            # SampleNonMP._elbo (reference SampleNonMP.py:56-57): logsumexp over the one K axis (no eps) - log K
            lf = self.plan_nonmp(None, self.P, self.Q, (), self.scope)
            if tuple(lf.axes) != (NONMP_K,):
                raise Exception(f"non-MP log-probability has axes {lf.axes}, expected ({NONMP_K},)")
            self.nonmp_lf = lf
            self.emit(ReduceOp(R_LSE, lp, [], [self.axdim(NONMP_K)], lf.tensors,
                               cadd=lf.const - math.log(self.sizes[NONMP_K]), tag='lp'))
        else:
            lf = self.plan_plate(None, self.P, self.Q, (), self.scope)
            if lf.axes != ():
                raise Exception(f"log-evidence has leftover axes {lf.axes}")
            root = self._last_contraction_as_root(lf)
            if root is not None:
                old = root.out
                root.out, root.cadd, root.tag = lp, root.cadd + lf.const, 'lp:' + root.tag
                self.producer[lp.id] = root
                if old.id in self.needs:
                    self.needs.add(lp.id)
            else:
                self.emit(ReduceOp(R_SUM, lp, [], [], lf.tensors, cadd=lf.const, tag='lp'))
        self.fwd_segments.append(self.fwd)
        plan = self.plan
        plan.programs = list(self.fwd_segments)
        plan.n_fwd = len(self.fwd_segments)
        bwd_segments = self.build_backward(grad_names)
        if self.fast_paths and os.environ.get("ALAN_B200_NO_SIDE") != "1":       # read when the plan is BUILT
            self.fuse_side_factors()
        plan.programs += bwd_segments
        plan.n_bwd = len(bwd_segments)
        if with_sample:
            plan.sample_prog = len(plan.programs)
            plan.programs.append(self.build_sampling_nonmp() if self.nonmp else self.build_sampling())
        plan.assign_offsets(self.itemsize)
        plan.serialize()
        return plan

    # -- backward -----------------------------------------------------------------------------
    def build_backward(self, grad_names):
        plan = self.plan
        plan.grad_inputs = list(grad_names)
        if not grad_names:
            return []
        all_fwd = []
        seg_of = {}
        sharded_ids = set()
        seg_ops = set(op for seg in self.fwd_segments for op in seg)
        for si, seg in enumerate(self.fwd_segments):
            for op in seg:
                lop = getattr(op, 'autodiff_as', None)
                if isinstance(lop, str):
                    continue
                lop = lop if lop is not None else op
                all_fwd.append(lop)
                seg_of[id(lop)] = si
                if si == 0:
                    sharded_ids.add(id(lop))
        needs = self.needs
        self.alloc_group = 1
        n_uses = {}
        for op in all_fwd:
            for x in self.op_inputs(op):
                n_uses[x.id] = n_uses.get(x.id, 0) + 1
        lazy_bcast = {}              # pt.id -> (small gout PT, its dims): adjoint = broadcast, never materialised
        adj = {}
        grad_out = {}
        for i, n in enumerate(grad_names):
            src = self.inputs[n]
            g = PT(src.axes, src.pos_shape, self.sizes, 'output', index=i, name=f'g:{n}')
            adj[src.id] = g
            grad_out[n] = g

        def adjoint(pt):
            if pt.id not in adj:
                adj[pt.id] = self.ws(pt.axes, pt.pos_shape, name=f'adj:{pt.name}')
            return adj[pt.id]

        segs = [[] for _ in self.fwd_segments]
        # seed: d lp / d lp_ws = upstream gradient (aux[0])
        seed = PT((), (), self.sizes, 'aux', index=0, name='grad_lp')
        adj[self.lp_ws.id] = seed
        sharded_region = set()
        if self.shard_plate is not None and len(self.fwd_segments) > 1:
            sharded_region = sharded_ids
        # Sharded plans: an adjoint that lacks the shard axis holds either a value replicated on every
        # rank (it only depends on the all-reduced tile) or a per-shard PARTIAL sum.  `partial_adj` =
        # tensors whose adjoint is partial: consumed, directly or transitively, inside the sharded
        # region.  User-visible global gradients are all-reduced once, so replicated contributions
        # into them (and into partial adjoints) are pre-divided by the world size (exact for 2/4/8).
        partial_adj = set()
        if self.shard_plate is not None:
            for op in reversed(all_fwd):
                if _iterates(op, self.shard_plate) or op.out.id in partial_adj:
                    for x in self.op_inputs(op):
                        if self.shard_plate not in x.axes:
                            partial_adj.add(x.id)

        def contribution_scale(op, target_pt, g):
            if self.shard_plate is None or self.shard_plate in target_pt.axes:
                return 1.0
            if not (g.space == 'output' or target_pt.id in partial_adj):
                return 1.0
            is_partial = _iterates(op, self.shard_plate) or op.out.id in partial_adj
            return 1.0 if is_partial else 1.0 / self.world_size

        for op in reversed(all_fwd):
            if op.out.space == 'output' and op.out is not self.lp_ws:
                continue
            if op.out.id not in needs or (op.out.id not in adj and op.out.id not in lazy_bcast):
                continue
            out_list = segs[len(self.fwd_segments) - 1 - seg_of[id(op)]]
            gout = adj.get(op.out.id)
            if isinstance(op, ExprOp) and self.fast_paths and getattr(op, 'fan_twin', None) is not None \
                    and op.fan_twin in seg_ops:
                self._fan_backward(op, op.fan_twin, gout, out_list, adjoint, needs, contribution_scale)
            elif isinstance(op, ExprOp):
                dims = op.keep + op.red
                for li, lf in enumerate(op.codeobj.leaves):
                    if lf.pt.id not in needs:
                        continue
                    g = adjoint(lf.pt)
                    kept = [d for d in dims if lf.stride(d) != 0]
                    kept.sort(key=lambda d: -lf.stride(d))
                    loop = [d for d in dims if lf.stride(d) == 0]
                    n_kept = _prod(d[2] for d in kept)
                    if n_kept != lf.pt.numel:
                        raise Exception(f"adjoint of {lf.pt}: expression does not cover the tensor")
                    n_loop = _prod(d[2] for d in loop)
                    nsplit = _choose_split(n_kept, n_loop)
                    scale = op.scale * contribution_scale(op, lf.pt, g)
                    if nsplit > 1:
                        part = self.ws_raw(nsplit * n_kept, name='partial_adj')
                        out_list.append(ExprBwdOp(part, op, li, kept, loop, gout, nsplit=nsplit, acc=0, scale=scale))
                        od = [('fl', 0, n_kept)]
                        out_list.append(ReduceOp(R_SUM, g, od, [('sp', 0, nsplit)],
                                                 [(_PartialRef(part, od, nsplit), 1.0)], acc=1))
                    else:
                        out_list.append(ExprBwdOp(g, op, li, kept, loop, gout, acc=1, scale=scale))
            elif isinstance(op, ReduceOp):
                dims = op.od + op.rd
                for fi, (lf, coeff) in enumerate(op.factors):
                    if lf.pt.id not in needs:
                        continue
                    if isinstance(lf, _PartialRef):
                        raise Exception("internal: partial buffers are never differentiated")
                    g = adjoint(lf.pt)
                    kept = [d for d in dims if lf.stride(d) != 0]
                    kept.sort(key=lambda d: -lf.stride(d))
                    loop = [d for d in dims if lf.stride(d) == 0]
                    n_kept = _prod(d[2] for d in kept)
                    if n_kept != lf.pt.numel:
                        raise Exception(f"adjoint of {lf.pt}: contraction does not cover the tensor")
                    n_loop = _prod(d[2] for d in loop)
                    nsplit = _choose_split(n_kept, n_loop)
                    scale = op.scale * coeff * contribution_scale(op, lf.pt, g)
                    if (op.mode == R_SUM and self.fast_paths and not loop and scale == 1.0 and n_uses.get(lf.pt.id) == 1
                            and isinstance(self.producer.get(lf.pt.id), FanLseOp) and g.space == 'ws'
                            and type(lf) is LeafRef and not lf.rename and not lf.mode):
                        # pure broadcast into the only consumer-less adjoint of a fused contraction: hand the
                        # small tensor to the adjoint kernel instead (it reads it through zero strides)
                        lazy_bcast[lf.pt.id] = (gout, op.od)
                        del adj[lf.pt.id]
                        continue
                    if op.mode == R_SUM:
                        mode, facs, kw = R_SUM, [(_OwnDims(gout, op.od), 1.0)], {}
                    else:
                        mode, facs = R_WSUM, op.factors
                        kw = dict(lse=(op.m_out, op.lo_out), gout=gout, lse_dims=op.od, gout_dims=op.od, cadd=op.cadd)
                    if nsplit > 1:
                        part = self.ws_raw(nsplit * n_kept, name='partial_adj')
                        out_list.append(ReduceOp(mode, part, kept, loop, facs, nsplit=nsplit, **kw))
                        od = [('fl', 0, n_kept)]
                        out_list.append(ReduceOp(R_SUM, g, od, [('sp', 0, nsplit)],
                                                 [(_PartialRef(part, od, nsplit), 1.0)], acc=1, scale=scale))
                    else:
                        out_list.append(ReduceOp(mode, g, kept, loop, facs, acc=1, scale=scale, **kw))
            elif isinstance(op, FanLseOp):
                bs = [(lf, coeff) for lf, coeff in op.bfactors if lf.pt.id in needs]
                qwant = op.qterm is not None and any(self._q_targets(op.qterm, needs))
                if bs or qwant:
                    if op.dense is not None:
                        lam, _, NG = op.dense
                        rows = [d for d in op.rho if d != lam] + [('sp', 0, NG), op.kappa]
                        gS = self.ws_raw(_prod(d[2] for d in rows), name='adj:small_factor_sum')
                    else:
                        rows = op.rho + [op.kappa]
                        gS = self.ws(tuple(d[1] for d in rows), name='adj:small_factor_sum')
                    if op.out.id in lazy_bcast:
                        gsmall, gdims = lazy_bcast[op.out.id]
                        out_list.append(FanLseBwdOp(op, gsmall, gS, gout_dims=gdims))
                    else:
                        out_list.append(FanLseBwdOp(op, gout, gS))
                    if qwant:
                        E, _, _, qc = op.qterm
                        if not self._try_normal_q_bwd(op, None, qc, gS, rows, out_list, adjoint, needs, n_uses,
                                                      contribution_scale, E=E):
                            raise Exception("internal: the inline Q factor has no fused adjoint for this shape")
                    for lf, coeff in bs:
                        if self.fast_paths and self._try_normal_q_bwd(op, lf, coeff, gS, rows, out_list, adjoint, needs,
                                                                        n_uses, contribution_scale):
                            continue
                        g = adjoint(lf.pt)
                        kept = [d for d in rows if lf.stride(d) != 0]
                        kept.sort(key=lambda d: -lf.stride(d))
                        loop = [d for d in rows if lf.stride(d) == 0]
                        n_kept = _prod(d[2] for d in kept)
                        if n_kept != lf.pt.numel:
                            raise Exception(f"adjoint of {lf.pt}: fused contraction does not cover the tensor")
                        scale = coeff * contribution_scale(op, lf.pt, g)
                        nsplit = _choose_split(n_kept, _prod(d[2] for d in loop))
                        facs = [(_OwnDims(gS, rows), 1.0)]
                        if nsplit > 1:
                            part = self.ws_raw(nsplit * n_kept, name='partial_adj')
                            out_list.append(ReduceOp(R_SUM, part, kept, loop, facs, nsplit=nsplit))
                            od = [('fl', 0, n_kept)]
                            out_list.append(ReduceOp(R_SUM, g, od, [('sp', 0, nsplit)],
                                                     [(_PartialRef(part, od, nsplit), 1.0)], acc=1, scale=scale))
                        else:
                            out_list.append(ReduceOp(R_SUM, g, kept, loop, facs, acc=1, scale=scale))
            elif isinstance(op, ChainOp):
                if op.ms.id in needs:
                    gms = adjoint(op.ms)
                    glevels = self.ws_raw(op.levels.numel, name='chain_glevels')
                    out_list.append(ChainBwdOp(op, gout, glevels, gms))
        # No zero-fill pass: every adjoint tensor is covered completely by each op that contributes to it (checked
        # above: "does not cover the tensor"), so its FIRST contribution -- first in execution order -- overwrites
        # (acc = 0) and only the later ones accumulate.  Gradient outputs that receive no contribution at all
        # (a parameter the log-evidence does not depend on) are the only tensors still zeroed explicitly.
        written = set()
        for seg in segs:
            for o in seg:
                if isinstance(o, NormalQBwdOp):
                    for attr, g in (('acc_l', o.g_l), ('acc_s', o.g_s)):
                        if g is not None:
                            setattr(o, attr, 1 if g.id in written else 0)
                            written.add(g.id)
                    continue
                dest = o.gleaf if isinstance(o, ExprBwdOp) else getattr(o, 'out', None)
                if dest is None or not isinstance(o, (ExprBwdOp, ReduceOp, ExprOp)):
                    continue
                if o.acc and dest.id not in written and not (isinstance(o, (ExprBwdOp, ReduceOp)) and o.nsplit > 1):
                    o.acc = 0
                written.add(dest.id)
        head = [FillOp(g, g.numel * self.itemsize) for n, g in grad_out.items() if g.id not in written]
        segs[0] = head + segs[0]
        if self.shard_plate is not None:
            plan.global_grads = [n for n in grad_names if self.shard_plate not in self.inputs[n].axes]
            if self.fused_collectives and plan.global_grads:
                pieces = [grad_out[n] for n in plan.global_grads]
                for k in range(0, len(pieces), 16):
                    segs[-1].append(XReduceOp(1 + k // 16, pieces[k:k + 16]))
        return segs

    def _q_targets(self, qterm, needs):
        """(loc wanted, scale / log-scale wanted) for an inline Q factor."""
        _, l, sc, _ = qterm
        A = self.producer.get(sc.pt.id)
        st = sc.pt
        if isinstance(A, ExprOp) and [ins[0] for ins in A.codeobj.instrs] == [VOPS['load'], VOPS['exp']]:
            st = A.codeobj.leaves[0].pt
        return l.pt.id in needs, st.id in needs

    def _try_normal_q_bwd(self, fanop, lf, coeff, gS, rows, out_list, adjoint, needs, n_uses, contribution_scale, E=None):
        """Recognise, among the small factors of a fused contraction, the mean-field Gaussian Q factor
        `sum_d log N(v[u,kappa,d]; loc[u,d], scale[u,d])` (scale possibly `exp` of a log-scale parameter) whose only
        consumer is this contraction, and emit its whole adjoint as ONE NormalQBwdOp that reads the contraction's
        adjoint gS directly -- instead of: sum gS over its partial slots, one gather-style ExprBwd per target leaf, the
        exp adjoint.  Returns False (nothing emitted) when the pattern does not apply."""
        if E is None:
            if type(lf) is not LeafRef or lf.rename or lf.mode or n_uses.get(lf.pt.id) != 1:
                return False
            E = self.producer.get(lf.pt.id)
        parts = self._normal3_parts(E)
        if parts is None:
            return False
        v, l, sc = parts
        D = E.red[0][2] if E.red else 1
        if D not in self.FAN_EVENT_EXTENTS:
            return False
        kappa = fanop.kappa
        ev = ('ev', 0, D)
        users = [d for d in E.keep if (d[0], d[1]) != (kappa[0], kappa[1])]
        if len(users) + 1 != len(E.keep) or v.stride(kappa) == 0 or v.pt.id in needs:
            return False
        n_users = _prod(d[2] for d in users)
        # the scale: a leaf of its own, or exp(log-scale leaf) hoisted into an 'arg' tensor that only this factor reads
        s_target, scale_is_exp = sc.pt, False
        A = self.producer.get(sc.pt.id)
        if isinstance(A, ExprOp) and [ins[0] for ins in A.codeobj.instrs] == [VOPS['load'], VOPS['exp']] \
                and not A.red and not A.acc and A.scale == 1.0 and n_uses.get(sc.pt.id, 0) <= 1:
            ls = A.codeobj.leaves[0]
            if type(ls) is LeafRef and not ls.rename and not ls.mode and ls.pt.shape == sc.pt.shape and ls.pt.axes == sc.pt.axes:
                s_target, scale_is_exp = ls.pt, True
        elif A is not None:
            return False
        for x in (l, sc):
            if x.stride(kappa) != 0 or x.pt.numel != n_users * D or any(x.stride(d) == 0 for d in users if d[2] > 1) \
                    or (D > 1 and x.stride(ev) == 0):
                return False
        want_l, want_s = l.pt.id in needs, s_target.id in needs
        if not (want_l or want_s):
            return False
        # gS dims: the users (same order), kappa, and at most one extra dim that is summed (partial slots / loc samples)
        keys = lambda ds: [(d[0], d[1]) for d in ds]
        extra = [d for d in rows if (d[0], d[1]) not in keys(users) and (d[0], d[1]) != (kappa[0], kappa[1])]
        if len(extra) > 1 or [k for k in keys(rows) if k in keys(users)] != keys(users):
            return False
        g_l = adjoint(l.pt) if want_l else None
        g_s = adjoint(s_target) if want_s else None
        c = coeff * contribution_scale(fanop, l.pt if want_l else s_target, g_l if want_l else g_s)
        op = NormalQBwdOp(D, users, kappa, v, l, sc, scale_is_exp, gS, rows, extra[0] if extra else None, c, g_l, g_s)
        op.gen = (E, A if scale_is_exp else None, lf)
        out_list.append(op)
        return True

    def _fan_backward(self, op, fan, gout, out_list, adjoint, needs, contribution_scale):
        """Adjoint of a materialised normal_fan factor through the fused kernels (reparameterised / VI
        gradients): R = d out / d(-v) per row, reduced by plain sums into the value and location adjoints;
        V and Wsum partials, reduced in a fixed order, give the scale adjoint."""
        D, F = fan.D, fan.F
        ev = ('ev', 0, D)
        rows = list(fan.rows)
        n_rows = _prod(d[2] for d in rows)
        dummy = self.ws_raw(1, name='fan_bwd_unused')
        want_rows = [(lf, sg) for lf, sg in ((fan.v, -1.0), (fan.l, 1.0)) if lf.pt.id in needs]
        if want_rows:
            R = self.ws_raw(n_rows * D, name='fan_R')
            out_list.append(NormalFanBwdOp(0, fan, gout, R, dummy, dummy, 1))
            own = rows + [ev]
            for lf, sg in want_rows:
                g = adjoint(lf.pt)
                kept = [d for d in own if lf.stride(d) != 0]
                kept.sort(key=lambda d: -lf.stride(d))
                loop = [d for d in own if lf.stride(d) == 0]
                n_kept = _prod(d[2] for d in kept)
                if n_kept != lf.pt.numel:
                    raise Exception(f"adjoint of {lf.pt}: fan factor does not cover the tensor")
                scale = sg * contribution_scale(op, lf.pt, g)
                nsplit = _choose_split(n_kept, _prod(d[2] for d in loop))
                facs = [(_OwnDims(R, own), 1.0)]
                if nsplit > 1:
                    part = self.ws_raw(nsplit * n_kept, name='partial_adj')
                    out_list.append(ReduceOp(R_SUM, part, kept, loop, facs, nsplit=nsplit))
                    od = [('fl', 0, n_kept)]
                    out_list.append(ReduceOp(R_SUM, g, od, [('sp', 0, nsplit)],
                                             [(_PartialRef(part, od, nsplit), 1.0)], acc=1, scale=scale))
                else:
                    out_list.append(ReduceOp(R_SUM, g, kept, loop, facs, acc=1, scale=scale))
        if fan.s.pt.id in needs:
            s = fan.s
            g = adjoint(s.pt)
            n_cta = int(max(1, min(296, n_rows // 512)))
            part_v = self.ws_raw(n_cta * F * D, name='fan_V_partial')
            part_w = self.ws_raw(n_cta * F, name='fan_W_partial')
            out_list.append(NormalFanBwdOp(1, fan, gout, dummy, part_v, part_w, n_cta))
            Vt = self.ws_raw(F * D, name='fan_V')
            Wt = self.ws_raw(F, name='fan_Wsum')
            for src, dst, n in ((part_v, Vt, F * D), (part_w, Wt, F)):
                od = [('fl', 0, n)]
                out_list.append(ReduceOp(R_SUM, dst, od, [('sp', 0, n_cta)], [(_PartialRef(src, od, n_cta), 1.0)]))
            fdim = ('ax', fan.fan_axis, F) if fan.fan_axis else ('ax', '__nofan', 1)
            all_dims = [fdim, ev]
            keep = [d for d in all_dims if s.stride(d) != 0]
            keep.sort(key=lambda d: -s.stride(d))
            red = [d for d in all_dims if s.stride(d) == 0 and d[2] > 1]
            # d out / d scale = T / scale^3 - 1 / scale, so  g_s += V / s^3 - Wsum / s
            L, MUL, DIV, SUB = VOPS['load'], VOPS['mul'], VOPS['div'], VOPS['sub']
            instrs = [(L, 0, 0, 0, 0, 0), (L, 1, 1, 0, 0, 0), (L, 2, 2, 0, 0, 0), (MUL, 3, 2, 2, 0, 0), (MUL, 4, 3, 2, 0, 0),
                      (DIV, 5, 0, 4, 0, 0), (DIV, 6, 1, 2, 0, 0), (SUB, 7, 5, 6, 0, 0)]
            code = Code(instrs, [], 7, [_OwnDims(Vt, [fdim, ev]), _OwnDims(Wt, [fdim]), s])
            out_list.append(ExprOp(g, keep, red, code, acc=1, scale=contribution_scale(op, s.pt, g), tag='fan_scale_adj'))

    # -- resampling (sample_logpq.py:17-107, reduce_Ks.py:35-83) ---------------------------------
    def build_sampling(self):
        if self.N is None:
            raise Exception("resampling program needs the number of posterior samples N")
        plan = self.plan
        plan.N = self.N
        sizes = dict(self.sizes)
        sizes['N'] = self.N
        idx_pt = {}
        plan.sample_groups = []
        for gi, g in enumerate(self.groups):
            axes = ('N',) + tuple(self.g2plates[g])
            idx_pt[Kname(g)] = (PT(axes, (), sizes, 'output', index=gi, name=f'idx:{g}'),
                                [('ax', a, sizes[a]) for a in axes])
            plan.sample_groups.append((g, tuple(self.g2plates[g])))
        ops = []
        sampled = set()

        def steps_for_sampling(level):
            """The reference re-runs collect_lps at sampling time on factors that were already gathered at the
            sampled parent indices (sample_logpq.py:75-81): every parent K axis has become the sample axis N, so
            the contraction ORDER -- and with it the order in which the level's Ks are drawn and the uniforms are
            consumed -- can differ from the forward pass.  Gathering commutes with the LSE over this level's Ks,
            so the same order is replayed here on the ungathered factors (parent Ks kept as batch axes; the
            sample kernel gathers), re-using the forward intermediates when the order is the same."""
            steps, _ = self.level_steps[level]
            lfs, Ks_here = self.level_factors[level]
            if not Ks_here:
                return steps
            seen = lambda axes: tuple(dict.fromkeys('N' if (a.startswith('K_') and a not in Ks_here) else a for a in axes))
            path_s = greedy_path([seen(lf.axes) for lf in lfs], Ks_here, sizes)
            if path_s == self.level_paths[level]:
                return steps
            out_steps = []
            cur = list(lfs)
            for idxs in path_s:
                chosen = [cur[i] for i in idxs]
                cur = [cur[i] for i in range(len(cur)) if i not in idxs]
                remaining = set(a for lf in cur for a in lf.axes)
                chosen_axes = _union_axes([lf.axes for lf in chosen])
                ks = tuple(k for k in Ks_here if k in chosen_axes and k not in remaining)
                tensors = [tc for lf in chosen for tc in lf.tensors]
                if not ks:
                    cur.append(LogicalFactor(tensors, 0.0, self.canon_order(chosen_axes)))
                    continue
                out_axes = self.canon_order([a for a in chosen_axes if a not in ks])
                out = self.ws(out_axes, name='lse_s[' + ','.join(ks) + ']')
                ops.append(ReduceOp(R_LSE_EPS, out, [self.axdim(a) for a in out_axes], [self.axdim(a) for a in ks],
                                    tensors, tag='contract_for_sampling:' + ','.join(ks)))
                out_steps.append(Step(tuple(level), tensors, ks))
                cur.append(LogicalFactor([(plain(out), 1.0)], 0.0, out_axes))
            return out_steps

        def visit(level, Q):
            steps = steps_for_sampling(level)
            for st in reversed(steps):
                batch_axes = tuple(a for a in self.all_plates
                                   if any(a in lf.pt.axes for lf, _ in st.tensors))
                if set(batch_axes) != set(level):
                    raise Exception("resampling a step whose factors do not carry every active plate is not supported")
                for k in st.ks:
                    g = k[2:]
                    if tuple(self.g2plates[g]) != tuple(batch_axes):
                        raise Exception(f"resampling: group {g} plates {self.g2plates[g]} != step plates {batch_axes}")
                batch = [('ax', 'N', self.N)] + [self.axdim(a) for a in batch_axes]
                idx_tensors, slot_of = [], {}
                facs = []
                for lf, coeff in st.tensors:
                    gathered = []
                    for a in lf.pt.axes:
                        if a.startswith('K_') and a not in st.ks:
                            if a not in sampled:
                                raise Exception(f"resampling order error: {a} needed before it was sampled")
                            if a not in slot_of:
                                slot_of[a] = len(idx_tensors)
                                idx_tensors.append(idx_pt[a])
                            gathered.append((a, slot_of[a]))
                    facs.append((lf, coeff, gathered))
                u_axes = batch_axes + ('N',)
                ui = len(plan.sample_steps)
                u = PT(u_axes, (), sizes, 'aux', index=ui, name=f'u{ui}')
                plan.sample_steps.append((batch_axes, st.ks))
                outs = [idx_pt[k][0] for k in st.ks]
                ops.append(SampleOp(batch, [self.axdim(k) for k in st.ks], facs, idx_tensors,
                                    (u, [('ax', a, sizes[a]) for a in u_axes]), outs))
                sampled.update(st.ks)
            for childname, childQ in Q.grouped_prog.items():
                if isinstance(childQ, Plate):
                    visit((*level, childname), childQ)
        visit((), self.Q)
        for g in self.groups:
            if Kname(g) not in sampled:
                raise Exception(f"importance_sample through a Timeseries is unfinished in the reference "
                                f"(README.md:41-44) and not provided here (group {g})")
        return ops


def weighted_moment_plan(x_sigs: dict, w_axes, f, sizes, dtype, canon) -> Plan:
    """Plan of `Marginals.moments` (reference Marginals.py:31-46 -> RawMoment.from_marginals, moments.py:16-35):

        out[plates..., *f.shape] = sum_{K axes of w}  f(x_1, ..., x_n) * w

    x_sigs: {varname: TensorSig} of the sample tensors the moment function takes (in argument order);
    w_axes: named axes of the marginal weights (K axes + plates).  One ExprOp (the traced f times w, nothing
    summed) into the workspace and one fixed-order ReduceOp over the K axes into output 0."""
    sig = dict(x_sigs)
    sig['__w'] = TensorSig('elf', tuple(w_axes), ())
    pl = Planner.bare(sig, sizes, dtype, canon)
    xs = [Expr.leaf(pl.inputs[v], s.axes, s.pos_shape) for v, s in x_sigs.items()]
    fx = trace_function(f, xs)
    kw = [a for a in w_axes if a.startswith('K_')]
    for a in fx.axes:
        if a.startswith('K_') and a not in kw:
            raise Exception(f"moment function depends on {a}, which the marginal weights do not carry")
    body = pl._prepare(Expr.make('mul', fx, Expr.leaf(pl.inputs['__w'], tuple(w_axes), ())))
    prod = pl.emit_expr(body, nred=0, tag='f*w')
    R = len(body.pos_shape)
    plates = [a for a in prod.axes if not a.startswith('K_')]
    ev = [('ev', R - 1 - i, int(body.pos_shape[i])) for i in range(R)]
    od = [pl.axdim(a) for a in plates] + ev
    rd = [pl.axdim(a) for a in prod.axes if a.startswith('K_')]
    out = PT(tuple(plates), body.pos_shape, pl.sizes, 'output', index=0, name='moment')
    pl.emit(ReduceOp(R_SUM, out, od, rd, [(plain(prod), 1.0)], tag='sum_K f*w'))
    plan = pl.plan
    plan.programs = [pl.fwd]
    plan.n_fwd, plan.n_bwd = 1, 0
    plan.out_axes, plan.out_shape = tuple(plates), tuple(out.shape)
    plan.assign_offsets(pl.itemsize)
    plan.serialize()
    return plan


class _PartialRef(LeafRef):
    """Reads a split-partial buffer laid out [nsplit, own dims...]."""
    def __init__(self, pt, own_dims, nsplit):
        object.__setattr__(self, 'pt', pt)
        object.__setattr__(self, 'rename', ())
        object.__setattr__(self, 'mode', 0)
        object.__setattr__(self, 'mdim', None)
        object.__setattr__(self, 'own', [('sp', 0, nsplit)] + list(own_dims))

    def stride(self, dim):
        st, acc = {}, 1
        for d in reversed(self.own):
            st[(d[0], d[1])] = acc if d[2] > 1 else 0
            acc *= d[2]
        return st.get((dim[0], dim[1]), 0)

    def __hash__(self):
        return id(self)

    def __eq__(self, o):
        return self is o


class _OwnDims(_PartialRef):
    """Reads a contiguous tensor through an explicit dim list (adjoint of a reduce output)."""
    def __init__(self, pt, own_dims):
        object.__setattr__(self, 'pt', pt)
        object.__setattr__(self, 'rename', ())
        object.__setattr__(self, 'mode', 0)
        object.__setattr__(self, 'mdim', None)
        object.__setattr__(self, 'own', list(own_dims))


def _iterates(op, plate):
    """True if the op walks the given plate axis, i.e. it belongs to the sharded region."""
    if isinstance(op, ExprOp):
        dims = op.keep + op.red
    elif isinstance(op, ReduceOp):
        dims = op.od + op.rd
    elif isinstance(op, FanLseOp):
        dims = op.rho + [op.kappa]
    elif isinstance(op, BernDotSumOp):
        dims = op.od + op.rd
    elif isinstance(op, NormalPolySumOp):
        dims = op.rows + op.kd + op.zd
    elif isinstance(op, ChainOp):
        return plate in op.ms.axes
    else:
        dims = []
    return any(d[0] == 'ax' and d[1] == plate for d in dims)


def _union(exprs):
    axes, shape = [], ()
    from .trace import _bshape
    for e in exprs:
        for a in e.axes:
            if a not in axes:
                axes.append(a)
        shape = _bshape(shape, e.pos_shape)
    return tuple(axes), shape


def _union_axes(list_of_axes):
    out = []
    for ax in list_of_axes:
        for a in ax:
            if a not in out:
                out.append(a)
    return tuple(out)


def _prod(it):
    n = 1
    for x in it:
        n *= int(x)
    return n


def _outputs_contiguous(factors, od, rd):
    """True when, in the largest factor, consecutive OUTPUT cells are adjacent in memory and the
    reduced axes are strided: then one thread per output (coalesced across threads) beats one warp
    per output (lanes striding over the reduced axis)."""
    if not factors or not od or not rd:
        return False
    big = max(factors, key=lambda fc: fc[0].pt.numel)[0]
    so = [big.stride(d) for d in od if d[2] > 1]
    sr = [big.stride(d) for d in rd if d[2] > 1]
    so = [x for x in so if x] or [0]
    sr = [x for x in sr if x] or [0]
    return min(so) != 0 and (min(sr) == 0 or min(so) < min(sr))


def _choose_split(n_out, n_red, target=148 * 2048):
    """Split a long reduction with few outputs across CTAs; partials are summed in a fixed order by a
    second launch, so the result does not depend on scheduling."""
    if n_red < 256 or n_out >= target:
        return 1
    want = max(1, target // max(n_out, 1))
    return int(max(1, min(want, n_red // 32, 1024)))
