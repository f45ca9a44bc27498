"""TEST INFRASTRUCTURE ONLY -- a slow torch-CPU interpreter of plan ops.

It executes the planner's op lists (alan_b200/plan.py) with the semantics the CUDA kernels
implement, so that planner logic (dims, strides, adjoint wiring, resampling order) is checked
against the oracle in the GPU-less build container.  Adjoint ops are evaluated with torch
autograd on the gathered operands, i.e. independently of the hand-written VM reverse sweep
that the CUDA kernels use.  It is never imported by the product.
"""
import math

import torch as t

from alan_b200 import plan as PL

UN = {6: lambda a: -a, 7: t.exp, 8: t.log, 9: t.sigmoid, 10: lambda a: a * a, 11: t.sqrt,
      12: lambda a: 1 / a, 13: t.nn.functional.softplus, 14: t.tanh, 15: t.abs, 16: t.log1p, 18: t.lgamma,
      19: lambda a: a, 21: t.cos, 22: t.sin}
BI = {2: t.add, 3: t.sub, 4: t.mul, 5: t.div, 17: t.pow, 20: lambda a, b: (a < b).to(a.dtype)}
HALF_LOG_2PI = 0.91893853320467274178


def _logsigmoid(x):
    return t.minimum(x, t.zeros_like(x)) - t.log1p(t.exp(-x.abs()))


def _density(op, r, a, b, c, d):
    if op == 32:
        dd = r[a] - r[b]
        return -(dd * dd) / (2 * (r[c] * r[c])) - t.log(r[c]) - HALF_LOG_2PI
    if op == 33:
        return -((1 - r[a]) * r[b] - _logsigmoid(r[b]))
    if op == 34:
        eps = t.finfo(r[b].dtype).eps
        pc = r[b].clamp(eps, 1 - eps)
        x = t.log(pc) - t.log1p(-pc)
        return -((1 - r[a]) * x - _logsigmoid(x))
    if op == 37:
        return t.log(r[b]) - r[b] * r[a]
    if op == 36:
        return -t.log(2 * r[c]) - (r[a] - r[b]).abs() / r[c]
    if op == 38:
        return t.xlogy(r[b], r[c]) + t.xlogy(r[b] - 1, r[a]) - r[c] * r[a] - t.lgamma(r[b])
    if op == 39:
        return (t.xlogy(r[b] - 1, r[a]) + t.xlogy(r[c] - 1, 1 - r[a]) + t.lgamma(r[b] + r[c])
                - t.lgamma(r[b]) - t.lgamma(r[c]))
    if op in (45, 46, 47, 48):
        x = r[c]
        if op in (46, 48):
            eps = t.finfo(x.dtype).eps
            pc = x.clamp(eps, 1 - eps)
            x = t.log(pc) - t.log1p(-pc)
        if op in (45, 46):
            return t.distributions.NegativeBinomial(r[b], logits=x, validate_args=False).log_prob(r[a])
        return t.distributions.Binomial(r[b], logits=x, validate_args=False).log_prob(r[a])
    if op == 40:
        return t.xlogy(r[a], r[b]) - r[b] - t.lgamma(r[a] + 1)
    if op == 35:
        lx = t.log(r[a]); dd = lx - r[b]
        return -(dd * dd) / (2 * (r[c] * r[c])) - t.log(r[c]) - HALF_LOG_2PI - lx
    raise NotImplementedError(f"emulator: density op {op}")


class Emu:
    def __init__(self, plan, inputs, outputs=None, aux=None):
        self.plan = plan
        self.dtype = plan.dtype
        self.item = 4 if self.dtype == t.float32 else 8
        self.ws = t.full((plan.ws_bytes // self.item + 64,), float('nan'), dtype=self.dtype)   # nothing may rely on a zeroed workspace
        self.inputs = [x.reshape(-1) for x in inputs]
        self.outputs = outputs or {}
        self.aux = aux or {}
        self.side = {}

    def buf(self, pt):
        if pt.space == 'ws':
            if pt.offset is None:             # tensor that only the unfused form of an op would touch
                if pt.id not in self.side:
                    self.side[pt.id] = t.full((pt.numel,), float('nan'), dtype=self.dtype)
                return self.side[pt.id], 0
            return self.ws, pt.offset // self.item
        if pt.space == 'input':
            return self.inputs[pt.index], 0
        if pt.space == 'output':
            return self.outputs[pt.index], 0
        if pt.space == 'aux':
            return self.aux[pt.index], 0
        raise Exception(pt.space)

    @staticmethod
    def grid(dims):
        sizes = [d[2] for d in dims]
        if not sizes:
            return []
        return list(t.meshgrid(*[t.arange(s) for s in sizes], indexing='ij'))

    def offsets(self, strides, grid):
        off = 0
        for s, g in zip(strides, grid):
            off = off + s * g
        if not grid:
            return t.zeros((), dtype=t.long)
        return off + t.zeros_like(grid[0])

    def load_leaf(self, lf, dims, grid):
        buf, base = self.buf(lf.pt)
        strides = [lf.stride(d) for d in dims]
        off = self.offsets(strides, grid) + base
        if lf.mode == 0:
            return buf[off], off, None
        k = [(d[0], d[1]) for d in dims].index(('ax', lf.mdim))
        if lf.mode == 1:
            mask = grid[k] >= 1
            off2 = t.where(mask, off - strides[k], t.full_like(off, base))
            return t.where(mask, buf[off2], t.zeros((), dtype=buf.dtype)), off2, mask
        mask = grid[k] == 0
        return t.where(mask, buf[off], t.zeros((), dtype=buf.dtype)), off, mask

    def vm(self, code, leafvals):
        r = {}
        for i, (op, dst, a, b, c, d) in enumerate(code.instrs):
            if op == 0:
                r[dst] = leafvals[a]
            elif op == 1:
                r[dst] = t.tensor(code.consts[a], dtype=self.dtype)
            elif op in UN:
                r[dst] = UN[op](r[a])
            elif op in BI:
                r[dst] = BI[op](r[a], r[b])
            else:
                r[dst] = _density(op, r, a, b, c, d)
        return r[code.res]

    # ---- ops
    def run(self, ops):
        for op in ops:
            getattr(self, 'op_' + type(op).__name__)(op)

    def op_ReduceSeqOp(self, op):
        """a run of small reductions in one launch: same results as the ops one by one"""
        self.run(op.ops)

    def op_DepsOp(self, op):
        """the dependency table of a program: every listed index precedes its op (a chain satisfies it)"""
        for k, d in enumerate(op.deps):
            assert all(0 < j < k for j in d), (k, d)

    def op_XReduceOp(self, op):
        """cross-rank sum in place (the real collective over gloo when a process group exists, identity otherwise)"""
        import torch.distributed as dist
        for pt in op.pieces:
            buf, base = self.buf(pt)
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                x = buf[base:base + pt.numel].clone()
                dist.all_reduce(x)
                buf[base:base + pt.numel] = x

    def op_PermOp(self, op):
        ub, ubase = self.buf(op.u)
        u = ub[ubase:ubase + op.rows * op.K].reshape(op.rows, op.K).double()
        if op.mode == 1:
            p = (u * op.K).long().clamp(max=op.K - 1)
        else:
            p = t.argsort(u, dim=-1, stable=True)
        self.side[('perm', op.out.id)] = p                      # int64: kept beside the (float) workspace

    def op_MvnPrepOp(self, op):
        """csrc/mvn.cuh with torch.linalg: scale_tril, its inverse and the log-normaliser per matrix."""
        sb, sbase = self.buf(op.S)
        d, n = op.d, op.n_mat
        if op.mode == 3:                                            # low rank: factor [n, d, r] + diagonal [n, d]
            Dg_pt, r = op.low_rank
            F = sb[sbase:sbase + n * d * r].reshape(n, d, r)
            db, dbase = self.buf(Dg_pt)
            S = F @ F.mT + t.diag_embed(db[dbase:dbase + n * d].reshape(n, d))
        else:
            S = sb[sbase:sbase + n * d * d].reshape(n, d, d)
        if op.mode in (0, 3):
            L = t.linalg.cholesky(S)
        elif op.mode == 2:
            L = S.tril()
        else:                                                       # torch's _precision_to_scale_tril
            Lf = t.linalg.cholesky(t.flip(S, (-2, -1)))
            L_inv = t.transpose(t.flip(Lf, (-2, -1)), -2, -1)
            L = t.linalg.solve_triangular(L_inv, t.eye(d, dtype=S.dtype), upper=False)
        W = t.linalg.solve_triangular(L, t.eye(d, dtype=S.dtype).expand(n, d, d), upper=False)
        c = -L.diagonal(dim1=-2, dim2=-1).log().sum(-1) - 0.5 * d * math.log(2 * math.pi)
        for pt, x in ((op.L, L), (op.W, W), (op.c, c)):
            b, base = self.buf(pt)
            b[base:base + x.numel()] = x.reshape(-1)

    def op_PasteOp(self, op):
        sb, sbase = self.buf(op.src)
        db, dbase = self.buf(op.dst)
        dims = [d for d in op.dims if d[0] > 1] or [(1, 0, 0)]
        grid = t.meshgrid(*[t.arange(d[0]) for d in dims], indexing='ij')
        so = sum(g * d[1] for g, d in zip(grid, dims)) + sbase
        do = sum(g * d[2] for g, d in zip(grid, dims)) + dbase
        db[do.reshape(-1)] = sb[so.reshape(-1)]

    def op_KGatherOp(self, op):
        xb, xbase = self.buf(op.x)
        x = xb[xbase:xbase + op.outer * op.K * op.inner].reshape(op.outer, op.K, op.inner)
        p = self.side[('perm', op.perm.id)]
        out = t.gather(x, 1, p.reshape(op.outer, op.K, 1).expand(op.outer, op.K, op.inner))
        ob, obase = self.buf(op.out)
        ob[obase:obase + out.numel()] = out.reshape(-1)

    def op_TsSampleOp(self, op):
        e = op.expr
        dims = e.keep
        ib, ibase = self.buf(op.init)
        KE = op.K * op.E
        prev = ib[ibase:ibase + op.n_outer * KE].reshape(op.n_outer, op.K, op.E).clone()
        perm = self.side[('perm', op.perm.id)].reshape(op.n_outer, op.T, op.K) if op.perm is not None else None
        ob, obase = self.buf(e.out)
        sizes = [d[2] for d in dims]
        grid = self.grid(dims)
        out = t.zeros(op.n_outer, op.T, op.K, op.E, dtype=self.dtype)
        for tt in range(op.T):
            vals = []
            for li, lf in enumerate(e.codeobj.leaves):
                if li == op.prev_leaf:
                    full = prev.reshape(op.n_outer, 1, op.K, op.E).expand(op.n_outer, op.T, op.K, op.E).reshape(sizes)
                    vals.append(full)
                else:
                    vals.append(self.load_leaf(lf, dims, grid)[0] + t.zeros(sizes, dtype=self.dtype))
            res = (self.vm(e.codeobj, vals) + t.zeros(sizes, dtype=self.dtype)).reshape(op.n_outer, op.T, op.K, op.E)
            cur = res[:, tt]
            out[:, tt] = cur
            if perm is not None:
                prev = t.gather(cur, 1, perm[:, tt].reshape(op.n_outer, op.K, 1).expand(op.n_outer, op.K, op.E))
            else:
                prev = cur
        ob[obase:obase + out.numel()] = out.reshape(-1)

    def op_FillOp(self, op):
        buf, base = self.buf(op.pt)
        n = op.nbytes // self.item
        buf[base:base + n] = 0

    def op_ExprOp(self, op):
        dims = op.keep + op.red
        grid = self.grid(dims)
        vals = [self.load_leaf(lf, dims, grid)[0] for lf in op.codeobj.leaves]
        res = self.vm(op.codeobj, vals)
        shape = [d[2] for d in dims]
        res = res + t.zeros(shape, dtype=self.dtype) if shape else res
        if op.red:
            res = res.sum(tuple(range(len(op.keep), len(dims))))
        buf, base = self.buf(op.out)
        n = res.numel()
        v = op.scale * res.reshape(-1)
        buf[base:base + n] = buf[base:base + n] + v if op.acc else v

    def op_FillRegionOp(self, op):
        lo, hi = self.plan.adj_region
        self.ws[lo // self.item: hi // self.item] = 0

    def op_FanLseOp(self, op):
        if op.qterm is not None:           # the fused kernel evaluates the Q factor inline; the emulator materialises it
            self.op_ExprOp(op.qterm[0])
        self.op_ExprOp(op.gen_expr)
        self.op_ReduceOp(op.gen_reduce)
        if op.psum is not None:
            # fused plate sum of the dense kernel: [rows, od...] partial rows; the emulator puts the whole sum over the
            # users into row 0 and zeros into the others
            part, rows, od = op.psum
            users = [d for d in op.rho if d != op.dense[0]]
            self.op_ReduceOp(PL.ReduceOp(PL.R_SUM, part, od, users, [(PL.plain(op.out), 1.0)]))
            buf, base = self.buf(part)
            n_out = math.prod(d[2] for d in od)
            buf[base + n_out: base + rows * n_out] = 0

    def op_FanLseBwdOp(self, op):
        f = op.fwd
        if f.qterm is not None:
            self.op_ExprOp(f.qterm[0])
        self.op_ExprOp(f.gen_expr)
        rows = f.rho + [f.kappa]
        fdim = [('ax', f.fan_axis, f.F)]
        r = f.gen_reduce
        if r.m_out is None:        # the fused kernel keeps (max, log-sum) in registers; the emulator materialises them
            r.m_out = PL.PT(r.out.axes, r.out.pos_shape, self.plan.sizes, 'ws', offset=None, name='emu_lse_m')
            r.lo_out = PL.PT(r.out.axes, r.out.pos_shape, self.plan.sizes, 'ws', offset=None, name='emu_lse_lo')
        saved = self.buf(r.out)[0][self.buf(r.out)[1]:self.buf(r.out)[1] + r.out.numel].clone()
        self.op_ReduceOp(r)        # recomputes out (identical) and fills the pair
        assert t.equal(self.buf(r.out)[0][self.buf(r.out)[1]:self.buf(r.out)[1] + r.out.numel], saved)
        kw = dict(lse=(r.m_out, r.lo_out), gout=op.gout, lse_dims=r.od,
                  gout_dims=op.gout_dims if op.gout_dims is not None else r.od, cadd=r.cadd)
        if f.dense is None:
            self.op_ReduceOp(PL.ReduceOp(PL.R_WSUM, op.gS, rows, fdim, r.factors, **kw))
            return
        # compact layout of the dense kernel (csrc/fan_tc2.cuh): gS[users, NG fan-group partials, kappa].  The
        # emulator puts the whole sum over (lam, f) into partial 0 and zeros into the others.
        lam, _, NG = f.dense
        users = [d for d in f.rho if d != lam] + [f.kappa]
        self.op_ReduceOp(PL.ReduceOp(PL.R_WSUM, op.gS, users, [lam] + fdim, r.factors, **kw))
        buf, base = self.buf(op.gS)
        n_u, Kk = math.prod(d[2] for d in users[:-1]), f.kappa[2]
        full = buf[base:base + n_u * Kk].clone().reshape(n_u, 1, Kk)
        out = t.cat([full, t.zeros(n_u, NG - 1, Kk, dtype=buf.dtype)], 1)
        buf[base:base + n_u * NG * Kk] = out.reshape(-1)

    def op_NormalFanBwdOp(self, op):
        f = op.fan
        fdim = ('ax', f.fan_axis, f.F) if f.fan_axis else ('ax', '__nofan', 1)
        ev = ('ev', 0, f.D)
        dims = list(f.rows) + [fdim, ev]
        grid = self.grid(dims)
        shape = [d[2] for d in dims]
        v = self.load_leaf(f.v, dims, grid)[0] + t.zeros(shape, dtype=self.dtype)
        l = self.load_leaf(f.l, dims, grid)[0] + t.zeros(shape, dtype=self.dtype)
        sc = self.load_leaf(f.s, dims, grid)[0] + t.zeros(shape, dtype=self.dtype)
        gb, gbase = self.buf(op.gout)
        o = PL.plain(f.out)
        G = gb[self.offsets([o.stride(d) if d[0] != 'ev' else 0 for d in dims], grid) + gbase]
        df = v - l
        nr = len(f.rows)
        if op.which == 0:
            R = (2 * df * G / (2 * sc * sc)).sum(nr)                      # sum over the fan axis -> [rows..., D]
            out, base = self.buf(op.R)
            out[base:base + R.numel()] = R.reshape(-1)
        else:
            rdims = tuple(range(nr))
            V = (G * df * df).sum(rdims) if rdims else G * df * df        # [F, D]
            Ws = G[..., 0].sum(rdims) if rdims else G[..., 0]              # [F]
            pv, bv = self.buf(op.partial)
            pw, bw = self.buf(op.partial_w)
            pv[bv:bv + op.n_cta * V.numel()] = 0
            pw[bw:bw + op.n_cta * Ws.numel()] = 0
            pv[bv:bv + V.numel()] = V.reshape(-1)
            pw[bw:bw + Ws.numel()] = Ws.reshape(-1)

    def op_NormalQBwdOp(self, op):
        ev = ('ev', 0, op.D)
        gdims = list(op.users) + ([op.sdim] if op.sdim is not None else []) + [op.kappa]
        ggrid = self.grid(gdims)
        gb, gbase = self.buf(op.gS)
        gref = PL._OwnDims(op.gS, op.gS_dims)
        G = gb[self.offsets([gref.stride(d) for d in gdims], ggrid) + gbase]
        if op.sdim is not None:
            G = G.sum(len(op.users))
        G = op.coeff * G                                                  # [users..., kappa]
        dims = list(op.users) + [op.kappa, ev]
        grid = self.grid(dims)
        shape = [d[2] for d in dims]
        v = self.load_leaf(op.v, dims, grid)[0] + t.zeros(shape, dtype=self.dtype)
        pdims = list(op.users) + [ev]
        pgrid = self.grid(pdims)
        lbuf, lbase = self.buf(op.l.pt)
        sbuf, sbase = self.buf(op.s.pt)
        loff = self.offsets([op.l.stride(d) for d in pdims], pgrid)
        soff = self.offsets([op.s.stride(d) for d in pdims], pgrid)
        l = lbuf[loff + lbase].clone().requires_grad_()
        sc = sbuf[soff + sbase].clone().requires_grad_()
        nu = len(op.users)
        with t.enable_grad():
            lp = (-((v - l.unsqueeze(nu)) ** 2) / (2 * sc.unsqueeze(nu) ** 2) - t.log(sc.unsqueeze(nu)) - HALF_LOG_2PI).sum(-1)
            gl, gs = t.autograd.grad((lp * G).sum(), [l, sc])
        if op.scale_is_exp:
            gs = gs * sc.detach()
        for g, val, off, acc in ((op.g_l, gl, loff, op.acc_l), (op.g_s, gs, soff, op.acc_s)):
            if g is None:
                continue
            b, base = self.buf(g)
            o = (off + base).reshape(-1)
            b[o] = b[o] + val.reshape(-1) if acc else val.reshape(-1)

    def op_BernDotSumOp(self, op):
        for g in op.gen_ops:
            getattr(self, 'op_' + type(g).__name__)(g)
        if op.side is not None:            # the Gaussian factor of the same rows, a second output of the fused kernel
            self.op_ExprOp(op.side[0])

    def op_NormalPolySumOp(self, op):
        """the fused formula of csrc/normal_poly.cuh evaluated from the op's polynomial (not from the ops it replaced)"""
        dims = op.rows + op.kd + op.zd
        grid = self.grid(dims)
        zv = [self.load_leaf(lf, dims, grid)[0] for lf in op.zleaves]
        kv = [self.load_leaf(lf, dims, grid)[0] for lf in op.kleaves]
        shape = [d[2] for d in dims]
        r = t.zeros(shape, dtype=self.dtype)
        for coeff, zs, ks in op.zterms + op.kterms:
            term = t.full(shape, coeff, dtype=self.dtype)
            for i in zs:
                term = term * zv[i]
            for i in ks:
                term = term * kv[i]
            r = r + term
        sc = kv[op.scale_leaf] if op.scale_leaf >= 0 else t.full(shape, op.scale_const, dtype=self.dtype)
        nz = dims[-1][2]
        sc0 = sc.select(-1, 0)
        val = op.cadd - (r * r).sum(-1) / (2 * sc0 * sc0) - nz * (sc0.log() + HALF_LOG_2PI)
        od = op.rows + op.kd
        ob, obase = self.buf(op.out)
        ostr = [PL.plain(op.out).stride(d) for d in od]
        off = self.offsets(ostr, self.grid(od)) + obase
        ob[off.reshape(-1)] = val.reshape(-1)

    def op_DotOp(self, op):
        self.op_ExprOp(op.autodiff_as)

    def op_NormalFanOp(self, op):
        self.op_ExprOp(op.autodiff_as)

    def op_ExprBwdOp(self, op):
        f = op.fwd
        dims = f.keep + f.red
        grid = self.grid(dims)
        loaded = [self.load_leaf(lf, dims, grid) for lf in f.codeobj.leaves]
        shape = [d[2] for d in dims]
        vals = []
        for i, (v, off, mask) in enumerate(loaded):
            v = (v + t.zeros(shape, dtype=self.dtype)).clone()
            if i == op.target:
                v.requires_grad_()
            vals.append(v)
        res = self.vm(f.codeobj, vals)
        res = res + t.zeros(shape, dtype=self.dtype)
        gbuf, gbase = self.buf(op.gout)
        gstr = PL._strides_like(op.gout, f.keep, dims)
        go = gbuf[self.offsets(gstr, grid) + gbase]
        if res.requires_grad:
            (g,) = t.autograd.grad((res * go).sum(), vals[op.target], allow_unused=True)
        else:
            g = None                 # piecewise-constant expression (comparisons only): the VM's derivative is zero
        if g is None:
            g = t.zeros(shape, dtype=self.dtype)
        _, off, mask = loaded[op.target]
        off = off + t.zeros(shape, dtype=t.long)
        if mask is not None:
            g = t.where(mask, g, t.zeros((), dtype=g.dtype))
        tb, tbase = self.buf(f.codeobj.leaves[op.target].pt)
        numel = f.codeobj.leaves[op.target].pt.numel
        acc = t.zeros(numel, dtype=self.dtype)
        acc.index_add_(0, (off - tbase).reshape(-1), g.reshape(-1))
        acc = op.scale * acc
        out, obase = self.buf(op.gleaf)
        if op.nsplit > 1:
            # partial layout [nsplit, n_kept]; emulate by putting everything in split 0
            out[obase:obase + op.nsplit * numel] = 0
            out[obase:obase + numel] = acc
        else:
            out[obase:obase + numel] = out[obase:obase + numel] + acc if op.acc else acc

    def op_ReduceOp(self, op):
        dims = op.od + op.rd
        grid = self.grid(dims)
        shape = [d[2] for d in dims]
        s = t.zeros(shape, dtype=self.dtype)
        for lf, coeff in op.factors:
            buf, base = self.buf(lf.pt)
            off = self.offsets([lf.stride(d) for d in dims], grid) + base
            s = s + coeff * buf[off]
        rdims = tuple(range(len(op.od), len(dims)))
        if op.mode == PL.R_SUM:
            res = s.sum(rdims) if rdims else s
        elif op.mode == PL.R_WSUM:
            m_pt, lo_pt = op.lse
            mb, mbase = self.buf(m_pt)
            lb, lbase = self.buf(lo_pt)
            gb, gbase = self.buf(op.gout)
            mm = mb[self.offsets(PL._strides_like(m_pt, op.lse_dims, dims), grid) + mbase]
            l = lb[self.offsets(PL._strides_like(lo_pt, op.lse_dims, dims), grid) + lbase]
            g = gb[self.offsets(PL._strides_like(op.gout, op.gout_dims, dims), grid) + gbase]
            w = g * t.exp((s - mm) - l)
            res = w.sum(rdims) if rdims else w
        else:
            m = s.amax(rdims, keepdim=True)
            a = t.exp(s - m).sum(rdims)
            eps = t.finfo(self.dtype).eps if op.mode == PL.R_LSE_EPS else 0.0
            res = t.log(a + eps) + m.reshape(a.shape)
            if op.m_out is not None:
                for pt, val in ((op.m_out, m.reshape(a.shape)), (op.lo_out, t.log(a + eps))):
                    b, base = self.buf(pt)
                    b[base:base + val.numel()] = val.reshape(-1)
        out, obase = self.buf(op.out)
        n = res.numel()
        if op.nsplit > 1:
            out[obase:obase + op.nsplit * n] = 0
            out[obase:obase + n] = res.reshape(-1)
            return
        v = op.scale * res.reshape(-1) + (0.0 if op.mode == PL.R_WSUM else op.cadd)
        out[obase:obase + n] = out[obase:obase + n] + v if op.acc else v

    def _chain(self, ms):
        from oracle.logpq_oracle import chain_logmmexp
        r = chain_logmmexp(ms.movedim(1, 0))            # [T, outer, K, K] -> [outer, K, K]
        return t.logsumexp(r, -1)

    def op_ChainOp(self, op):
        buf, base = self.buf(op.ms)
        n = op.outer * op.T * op.K * op.K
        ms = buf[base:base + n].reshape(op.outer, op.T, op.K, op.K)
        out, obase = self.buf(op.out)
        out[obase:obase + op.outer * op.K] = self._chain(ms).reshape(-1)

    def op_ChainBwdOp(self, op):
        f = op.fwd
        buf, base = self.buf(f.ms)
        n = f.outer * f.T * f.K * f.K
        ms = buf[base:base + n].reshape(f.outer, f.T, f.K, f.K).clone().requires_grad_()
        gb, gbase = self.buf(op.gout)
        go = gb[gbase:gbase + f.outer * f.K].reshape(f.outer, f.K)
        (g,) = t.autograd.grad((self._chain(ms) * go).sum(), ms)
        out, obase = self.buf(op.gms)
        out[obase:obase + n] = g.reshape(-1)

    def op_SampleOp(self, op):
        grid = self.grid(op.batch)
        shape = [d[2] for d in op.batch]
        gv = []
        for pt, own in op.idx_tensors:
            b, base = self.buf(pt)
            gv.append(b[self.offsets(PL._strides_like(pt, own, op.batch), grid) + base])
        ksz = [d[2] for d in op.ks]
        ktot = math.prod(ksz)
        kgrid = t.meshgrid(*[t.arange(s) for s in ksz], indexing='ij')
        lp = t.zeros(shape + [ktot], dtype=self.dtype)
        for lf, coeff, gathered in op.factors:
            b, base = self.buf(lf.pt)
            off = self.offsets([lf.stride(d) for d in op.batch], grid) + base
            for axis, slot in gathered:
                off = off + gv[slot] * lf.stride(('ax', axis, 0))
            koff = sum(lf.stride(d) * g for d, g in zip(op.ks, kgrid)).reshape(-1)
            lp = lp + coeff * b[off.unsqueeze(-1) + koff]
        pt, own = op.u
        ub, ubase = self.buf(pt)
        u = ub[self.offsets(PL._strides_like(pt, own, op.batch), grid) + ubase]
        m = lp.amax(-1, keepdim=True)
        p = t.exp(lp - m).double()                                 # factor dtype, then float64 (Appendix A8)
        c = p.cumsum(-1)
        thr = (u * c[..., -1]).unsqueeze(-1)
        pick = (c < thr).sum(-1).clamp(max=ktot - 1)
        for k in range(len(ksz) - 1, -1, -1):
            o, obase = self.buf(op.outs[k])
            o[obase:obase + pick.numel()] = (pick % ksz[k]).reshape(-1)
            pick = pick // ksz[k]


class EmuRunner:
    """TEST INFRASTRUCTURE ONLY: the surface of engine.Runner (device_inputs / forward_raw / backward_raw / elbo)
    executed by the CPU emulator, so that the host-side glue above the C ABI (alan_adapter.B200, the Sample._elbo
    hook, autograd wiring) can be exercised in the GPU-less build container.  Never imported by the product."""
    def __init__(self, comp, device=None, process_group=None):
        self.comp, self.dtype, self.device = comp, comp.dtype, t.device("cpu")
        self.generation = 0
        self.emu = None

    def device_inputs(self, sample, inputs_params, data, extra_log_factors=None, differentiable=False):
        comp = self.comp
        src = {}
        for d in (sample, inputs_params or {}, data or {}):
            src.update(d)
        elf = dict(extra_log_factors or {})
        out = []
        for name in comp.plan.input_names:
            if name in comp.plan.const_inputs:
                x = comp.plan.const_inputs[name]
            elif name.startswith('__J'):
                _, plates, pos = next(m for m in comp.moment_inputs if m[0] == name)
                x = t.zeros([comp.sizes[a] for a in plates] + list(pos), dtype=self.dtype)
            else:
                key, role, orig, axes = next(o for o in comp.order if o[0] == name)
                v = elf[orig] if role == 'elf' else src[orig]
                x = v.order(axes).t.to(self.dtype).contiguous()
                if not differentiable or name not in comp.grad_names:
                    x = x.detach()
            out.append(x)
        return out

    def forward_raw(self, tensors):
        plan = self.comp.plan
        self.generation += 1
        lp = t.zeros(1, dtype=self.dtype)
        self.emu = Emu(plan, [x.detach() for x in tensors], outputs={0: lp})
        for seg in plan.programs[:plan.n_fwd]:
            self.emu.run(seg)
        return lp[0].clone()

    def backward_raw(self, tensors, grad_lp=None):
        plan = self.comp.plan
        g = t.ones(1, dtype=self.dtype) if grad_lp is None else grad_lp.detach().reshape(1).to(self.dtype)
        self.emu.outputs = {i: t.zeros(plan.input_pts[n].numel, dtype=self.dtype) for i, n in enumerate(plan.grad_inputs)}
        self.emu.aux = {0: g}
        with t.enable_grad():                  # the emulator evaluates adjoint ops with autograd on gathered operands
            for seg in plan.programs[plan.n_fwd:plan.n_fwd + plan.n_bwd]:
                self.emu.run(seg)
        return {n: self.emu.outputs[i].detach().reshape(plan.input_pts[n].shape) for i, n in enumerate(plan.grad_inputs)}

    def elbo(self, tensors):
        from alan_b200.engine import _LogPQFunction
        return _LogPQFunction.apply(self, *tensors)
