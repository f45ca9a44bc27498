"""Host-side mirror of the reference's L4->L3 seam for the logPQ path.

`LogPQ` plays the role of `Sample._elbo / _marginal_idxs / _moments_uniform_input /
_importance_sample_idxs / index_into_sample` (reference: src/alan/Sample.py:69-108, 150-183,
208-272, 291-346, 359-381) for a fixed model and tensor signature: it compiles the plan once
and then every call is a handful of C-ABI launches on the caller's CUDA stream.

There is no CPU fallback: constructing a `LogPQ` on a machine without the CUDA library and a
CUDA device raises.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from .model import Plate, Kname, check_PQ
from .named import NT
from .plan import Planner, TensorSig, Plan, NONMP_K
from .trace import Expr, trace_function


def working_dtype(*dicts):
    """fp64 if any floating input is fp64, else fp32 (the reference's promotion on this path,
    SURVEY.md §7 'dtype promotion')."""
    dt = torch.float32
    for d in dicts:
        for v in d.values():
            if v.t.dtype == torch.float64:
                dt = torch.float64
    return dt


NARROW_MIN_NUMEL = 1 << 18          # uint8 / bool host tensors of at least this many elements cross PCIe as bytes


def _to_working(t, dtype, device=None):
    """Move to `device` (if given) and cast to the working dtype.  uint8 / bool tensors cross PCIe as bytes and are
    widened on the device by the library's own kernel (alan_b200_widen_u8), not as four-byte floats."""
    from . import runtime
    if t.dtype == dtype and t.is_contiguous() and (device is None or t.device == device):
        return t                                     # already canonical: the common case in a training loop
    if device is not None:
        t = t.to(device)
    if t.dtype in runtime.NARROW_DTYPES and t.is_cuda:
        return runtime.widen(t.contiguous(), dtype=dtype)
    return t.to(dtype).contiguous()


class Compiled:
    """Plan + the canonical (contiguous, canonical axis order, working dtype) input list."""
    def __init__(self, P: Plate, Q: Plate, sample, inputs_params, data, extra_log_factors=None,
                 moment_specs=(), grad_names=(), N=None, shard_plate=None, world_size=1, dtype=None,
                 fast_paths=True, fused_collectives=False, nonmp=False):
        sample, inputs_params, data = dict(sample), dict(inputs_params or {}), dict(data or {})
        elf = dict(extra_log_factors or {})
        check_PQ(P, Q, set(data.keys()))
        self.P, self.Q = P, Q
        self.dtype = dtype or working_dtype(sample, inputs_params, data, elf)
        all_plates = P.all_platenames()
        groups = Q.groupvarnames()
        canon = list(all_plates) + ([NONMP_K] if nonmp else [Kname(g) for g in groups])
        sizes = {}
        named = {}
        for role, d in (('sample', sample), ('param', inputs_params), ('data', data), ('elf', elf)):
            for k, v in d.items():
                key = k if role != 'elf' else f"__elf{len([n for n in named if n.startswith('__elf')])}"
                if key in named:
                    raise Exception(f"name {key} is used twice among samples / inputs / params / data")
                for a, s in v.named_sizes.items():
                    if a not in canon:
                        raise Exception(f"{k}: axis {a} is neither a plate nor a K axis of this model")
                    if sizes.setdefault(a, s) != s:
                        raise Exception(f"{k}: axis {a} has size {s}, elsewhere {sizes[a]}")
                named[key] = (role, k, v)
        self.sizes = sizes
        self.canon = canon
        self.order_by_name = {}
        sig, self.order, self.elf_keys = {}, [], {}
        extra = []
        for key, (role, orig, v) in named.items():
            axes = tuple(a for a in canon if a in v.axes)
            sig[key] = TensorSig(role, axes, v.pos_shape, requires_grad=False)
            self.order.append((key, role, orig, axes))
            self.order_by_name[key] = (key, role, orig, axes)
            if role == 'elf':
                self.elf_keys[orig] = key
        planner = Planner(P, Q, sig, sizes, self.dtype, want_sample_N=N, shard_plate=shard_plate,
                          world_size=world_size, fast_paths=fast_paths,
                          fused_collectives=fused_collectives and world_size > 1 and N is None, nonmp=nonmp)
        for orig, key in self.elf_keys.items():
            s = sig[key]
            extra.append((orig, Expr.leaf(planner.inputs[key], s.axes, s.pos_shape)))
        # moments: factor  sum_pos f(x) * J   (Sample.py:326-338) with J a zero source term
        self.moment_inputs = []
        for i, (varnames, f) in enumerate(moment_specs):
            xs = [Expr.leaf(planner.inputs[v], sig[v].axes, sig[v].pos_shape) for v in varnames]
            fx = trace_function(f, xs)
            plates = tuple(a for a in all_plates if a in fx.axes)
            jname = f"__J{i}"
            jpt = planner._add_input(jname, plates, fx.pos_shape)
            sig[jname] = TensorSig('elf', plates, fx.pos_shape)
            self.moment_inputs.append((jname, plates, tuple(fx.pos_shape)))
            extra.append((jname, Expr.make('mul', fx, Expr.leaf(jpt, plates, fx.pos_shape))))
        planner.extra_factors = extra
        gnames = [self._key_of(n) for n in grad_names] + [j for j, _, _ in self.moment_inputs]
        self.grad_names = gnames
        planner.set_grad_names(gnames)
        self.plan: Plan = planner.build(grad_names=gnames, with_sample=N is not None)
        self.planner = planner

    def _key_of(self, name):
        if name in self.elf_keys:
            return self.elf_keys[name]
        return name

    def canonical_inputs(self, sample, inputs_params, data, extra_log_factors=None, device=None, keep_narrow=False):
        """Permute every tensor to canonical axis order, make it contiguous in the working dtype.  keep_narrow=True
        leaves uint8 / bool tensors in their one-byte type (host staging of PipelinedRunner: they are widened on the
        device after the copy)."""
        src = {}
        for d in (sample, inputs_params or {}, data or {}):
            src.update(d)
        elf = dict(extra_log_factors or {})
        out = []
        for name in self.plan.input_names:
            if name in self.plan.const_inputs:
                t = self.plan.const_inputs[name]
            elif name.startswith('__J'):
                _, plates, pos = next(m for m in self.moment_inputs if m[0] == name)
                t = torch.zeros([self.sizes[a] for a in plates] + list(pos), dtype=self.dtype)
            else:
                key, role, orig, axes = self.order_by_name[name]
                v = elf[orig] if role == 'elf' else src[orig]
                t = v.order(axes).t
            t = t.detach()
            # (small tensors are widened here, on the host: below ~1 MB saved the extra copy + widen launch costs more
            # than the bytes -- cfg-2's 27 KB of covariates: 0.153 -> 0.186 ms per pipelined step when kept narrow)
            if keep_narrow and t.dtype in (torch.uint8, torch.bool) and t.numel() >= NARROW_MIN_NUMEL:
                out.append((t.to(device) if device is not None else t).contiguous())
            else:
                out.append(_to_working(t, self.dtype, device))
        return out


class _LogPQFunction(torch.autograd.Function):
    """lp = logPQ(inputs); backward calls the hand-written adjoint program (no autograd
    re-materialisation, reference logpq.py:62-66)."""

    @staticmethod
    def forward(ctx, runner, *tensors):
        lp = runner.forward_raw(list(tensors))
        ctx.runner = runner
        ctx.generation = runner.generation
        ctx.save_for_backward(*tensors)
        return lp

    @staticmethod
    def backward(ctx, grad_lp):
        runner = ctx.runner
        tensors = list(ctx.saved_tensors)
        if runner.generation != ctx.generation:
            # another forward ran on this runner's workspace since ours (runners are shared by every Sample of a
            # Problem with the same signature; `(l1 + l2).backward()` is legal): the adjoint program reads forward
            # intermediates from the workspace, so recompute ours first -- what autograd's saved tensors give the
            # reference for free
            runner.forward_raw(tensors)
        grads = runner.backward_raw(tensors, grad_lp)
        out = [None] * len(tensors)
        for name, g in grads.items():
            out[runner.comp.plan.input_names.index(name)] = g
        return (None, *out)


class Runner:
    """One compiled plan on one device."""
    def __init__(self, comp: Compiled, device=None, process_group=None):
        from . import runtime
        self.comp = comp
        self.dp = runtime.DevicePlan(comp.plan, device)
        self.device = self.dp.device
        self.dtype = comp.dtype
        self.pg = process_group
        self.lp = torch.zeros((), dtype=self.dtype, device=self.device)
        self.generation = 0          # bumped by every forward: identifies whose intermediates the workspace holds
        if comp.plan.fused_collectives:
            # the cross-rank sums run inside the programs over NVLink peer memory: map the ranks' symmetric buffers
            self.dp.attach_symmetric(process_group)

    # ---- raw calls on canonical device tensors ------------------------------------------
    def forward_raw(self, tensors):
        plan = self.comp.plan
        self.generation += 1
        lp = torch.empty((), dtype=self.dtype, device=self.device)
        for seg in range(plan.n_fwd):
            self.dp.fwd(seg, tensors, lp)
            if seg + 1 < plan.n_fwd:
                self._allreduce_tile()
        return lp

    def _allreduce_tile(self):
        import torch.distributed as dist
        tile = self.dp.ws_view(self.comp.plan.allreduce, self.dtype)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.pg) > 1:
            dist.all_reduce(tile, group=self.pg)

    def backward_raw(self, tensors, grad_lp=None):
        plan = self.comp.plan
        if grad_lp is None:
            grad_lp = torch.ones((), dtype=self.dtype, device=self.device)
        grad_lp = grad_lp.to(self.dtype).contiguous()
        import torch.distributed as dist
        sharded = bool(plan.global_grads) and dist.is_available() and dist.is_initialized() \
            and dist.get_world_size(self.pg) > 1 and not plan.fused_collectives
        flat, views = None, {}
        if sharded:
            # the global-parameter gradients live side by side in ONE buffer (slices aligned to 16 bytes), so the
            # single all-reduce runs in place: no concatenate / copy-back launches around the collective
            step = 16 // torch.empty((), dtype=self.dtype).element_size()
            offs, total = {}, 0
            for n in plan.global_grads:
                offs[n] = total
                total += -(-plan.input_pts[n].numel // step) * step
            flat = torch.zeros(max(total, 1), dtype=self.dtype, device=self.device)
            for n in plan.global_grads:
                pt = plan.input_pts[n]
                views[n] = flat[offs[n]:offs[n] + pt.numel].view(pt.shape)
        outs = [views[n] if n in views else torch.empty(plan.input_pts[n].shape, dtype=self.dtype, device=self.device)
                for n in plan.grad_inputs]
        for seg in range(plan.n_bwd):
            self.dp.bwd(seg, tensors, grad_lp, outs)
        if sharded:
            dist.all_reduce(flat, group=self.pg)
        return dict(zip(plan.grad_inputs, outs))

    def step(self, tensors):
        """forward_raw + backward_raw as ONE replayed CUDA graph (the collectives of a sharded plan included):
        a training loop that keeps its device buffers pays one graph launch per step instead of one per
        program segment plus the eager collectives between them.  The first call with a given binding of
        input pointers runs eagerly, the second is captured, later ones replay.  Returns (lp, {name: grad});
        both are STATIC buffers that the next call with the same binding overwrites.  Any failure to
        capture (e.g. a collective backend that cannot be captured) falls back to the eager path for good."""
        if not hasattr(self, "_step_graphs"):
            self._step_graphs, self._step_seen, self._step_off = {}, {}, False
        key = tuple(int(x.data_ptr()) for x in tensors)
        ent = self._step_graphs.get(key)
        if ent is not None:
            self.generation += 1
            ent[0].replay()
            return ent[1], ent[2]
        if self._step_off or self._step_seen.get(key, 0) < 1 or len(self._step_graphs) >= 4:
            if len(self._step_seen) > 64:                    # a caller that never repeats a binding: keep this bounded
                self._step_seen.clear()
            self._step_seen[key] = self._step_seen.get(key, 0) + 1
            lp = self.forward_raw(tensors)
            return lp, self.backward_raw(tensors)
        try:
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                lp = self.forward_raw(tensors)
                grads = self.backward_raw(tensors)
            g.replay()
        except Exception:
            self._step_off = True
            torch.cuda.synchronize(self.device)
            lp = self.forward_raw(tensors)
            return lp, self.backward_raw(tensors)
        self._step_graphs[key] = (g, lp, grads)
        return lp, grads

    def resample_raw(self, tensors, uniforms):
        """uniforms: list of float64 device tensors, one per sampling step (plan.sample_steps order),
        each laid out [batch plates..., N].  Returns {groupvarname: int64 [N, plates...]}."""
        plan = self.comp.plan
        if plan.sample_prog < 0:
            raise Exception("this plan was compiled without a resampling program (pass N=...)")
        if len(uniforms) != len(plan.sample_steps):
            raise Exception(f"expected {len(plan.sample_steps)} uniform tensors, got {len(uniforms)}")
        us = []
        for u, (batch, ks) in zip(uniforms, plan.sample_steps):
            shape = [self.comp.sizes[a] for a in batch] + [plan.N]
            if list(u.shape) != shape or u.dtype != torch.float64:
                raise Exception(f"uniforms for step over {ks} must be float64 of shape {shape}")
            us.append(u.to(self.device).contiguous())
        outs = [torch.empty([plan.N] + [self.comp.sizes[a] for a in plates], dtype=torch.int64, device=self.device)
                for _, plates in plan.sample_groups]
        self.dp.resample(tensors, us, outs)
        return {g: NT(o, ('N',) + tuple(plates)) for (g, plates), o in zip(plan.sample_groups, outs)}

    # ---- named-tensor calls -----------------------------------------------------------------
    def device_inputs(self, sample, inputs_params, data, extra_log_factors=None, differentiable=False):
        """Canonical device tensors.  With differentiable=True the permute/cast is left on the
        autograd tape so that gradients flow back to the caller's tensors."""
        comp = self.comp
        if not differentiable:
            return comp.canonical_inputs(sample, inputs_params, data, extra_log_factors, device=self.device)
        src = {}
        for d in (sample, inputs_params or {}, data or {}):
            src.update(d)
        elf = dict(extra_log_factors or {})
        out = []
        for name in comp.plan.input_names:
            if name in comp.plan.const_inputs:
                t = self.dp.consts[name]
            elif name.startswith('__J'):
                _, plates, pos = next(m for m in comp.moment_inputs if m[0] == name)
                t = torch.zeros([comp.sizes[a] for a in plates] + list(pos), dtype=self.dtype, device=self.device)
                t.requires_grad_(True)
            else:
                key, role, orig, axes = comp.order_by_name[name]
                v = elf[orig] if role == 'elf' else src[orig]
                t = _to_working(v.order(axes).t, self.dtype, self.device)
                if name not in comp.grad_names:
                    t = t.detach()
            out.append(t)
        return out

    def elbo(self, tensors):
        """Differentiable log-evidence estimate on canonical device tensors."""
        return _LogPQFunction.apply(self, *tensors)


class SplitRunner:
    """`computation_strategy=Split(plate, n)` on one GPU (reference Split.py:44-130, logpq.py:43-57,151-153).

    The plate is processed in blocks of `n` elements (sizes by the reference's rule, strategy.Split.sizes) through
    ONE block-sized workspace per distinct block size: pass 1 runs the sharded part of the forward program for every
    block and adds the `[K_parents]` tiles left to right; the replicated top level then gives the log-evidence.
    The adjoint pass recomputes a block's forward before its backward (the reference wraps every chunk in
    torch.utils.checkpoint, logpq.py:62-66), so the device memory in use is that of one block.  Per-element
    gradients are written into slices of full-size outputs, global-parameter gradients are summed over the blocks.
    Same `forward_raw / backward_raw / comp` surface as `Runner`, so the autograd wrapper works with either."""
    def __init__(self, P, Q, sample, inputs_params, data, plate, split, extra_log_factors=None, moment_specs=(),
                 grad_names=(), device=None):
        self.plate = plate
        full = {}
        for d in (sample, inputs_params or {}, data or {}, extra_log_factors or {}):
            full.update({k: v for k, v in d.items()})
        M = None
        for v in full.values():
            if plate in v.axes:
                M = v.named_sizes[plate]
        if M is None:
            raise Exception(f"Split: no tensor carries the plate {plate}")
        self.M = M
        self.block_sizes = split.sizes(M)
        self.blocks, lo = [], 0
        for m in self.block_sizes:
            self.blocks.append((lo, lo + m))
            lo += m

        def head(d, m):
            out = {}
            for k, v in (d or {}).items():
                if plate in v.axes:
                    out[k] = NT(v.t.narrow(v.axes.index(plate), 0, m), v.axes)
                else:
                    out[k] = v
            return out
        self.runs = {}
        for m in dict.fromkeys(self.block_sizes):
            comp = Compiled(P, Q, head(sample, m), head(inputs_params, m), head(data, m),
                            extra_log_factors=head(extra_log_factors, m), moment_specs=moment_specs,
                            grad_names=list(grad_names), shard_plate=plate, world_size=len(self.block_sizes))
            if comp.plan.n_fwd != 2:
                raise Exception(f"Split: {plate} is not a top-level plate of this model (only those can be split here)")
            self.runs[m] = Runner(comp, device)
        self.comp = next(iter(self.runs.values())).comp          # names / order / dtype are those of any block plan
        self.device = next(iter(self.runs.values())).device
        self.dtype = self.comp.dtype
        self.generation = 0
        plan = self.comp.plan
        self.carries = []
        for name in plan.input_names:
            pt = plan.input_pts[name]
            if plate in pt.axes and pt.axes[0] != plate:
                raise Exception(f"Split: {name} carries {plate} but not as its outermost axis")
            self.carries.append(plate in pt.axes)
        self.one = torch.ones((), dtype=self.dtype, device=self.device)

    def device_inputs(self, sample, inputs_params, data, extra_log_factors=None, differentiable=False):
        """FULL-SIZE canonical device tensors in plan input order (cf. Runner.device_inputs)."""
        comp, M = self.comp, self.M
        src = {}
        for d in (sample, inputs_params or {}, data or {}):
            src.update(d)
        elf = dict(extra_log_factors or {})
        out = []
        run0 = next(iter(self.runs.values()))
        for name in comp.plan.input_names:
            if name in comp.plan.const_inputs:
                t = run0.dp.consts[name]
            elif name.startswith('__J'):
                _, plates, pos = next(m for m in comp.moment_inputs if m[0] == name)
                t = torch.zeros([(M if a == self.plate else comp.sizes[a]) for a in plates] + list(pos),
                                dtype=self.dtype, device=self.device)
                if differentiable:
                    t.requires_grad_(True)
            else:
                key, role, orig, axes = next(o for o in comp.order if o[0] == name)
                v = elf[orig] if role == 'elf' else src[orig]
                t = v.order(axes).t.to(self.device).to(self.dtype).contiguous()
                if not differentiable or name not in comp.grad_names:
                    t = t.detach()
            out.append(t)
        return out

    def elbo(self, tensors):
        """Differentiable log-evidence estimate on FULL-SIZE canonical device tensors."""
        return _LogPQFunction.apply(self, *tensors)

    def _block(self, tensors, b):
        lo, hi = self.blocks[b]
        return [x[lo:hi] if c else x for x, c in zip(tensors, self.carries)], self.runs[hi - lo]

    def forward_raw(self, tensors):
        self.generation += 1
        plan = self.comp.plan
        lp = torch.empty((), dtype=self.dtype, device=self.device)
        total = None
        for b in range(len(self.blocks)):
            tens, run = self._block(tensors, b)
            run.dp.fwd(0, tens, lp)
            tile = run.dp.ws_view(run.comp.plan.allreduce, self.dtype)
            total = tile.clone() if total is None else total + tile            # prev_lpq + lp, left to right
        self.tile = total
        tens, run = self._block(tensors, len(self.blocks) - 1)
        run.dp.ws_view(run.comp.plan.allreduce, self.dtype).copy_(total)
        run.dp.fwd(1, tens, lp)
        return lp

    def backward_raw(self, tensors, grad_lp=None):
        plan = self.comp.plan
        if grad_lp is None:
            grad_lp = self.one
        grad_lp = grad_lp.to(self.dtype).contiguous()
        lp = torch.empty((), dtype=self.dtype, device=self.device)
        full = {}
        for n in plan.grad_inputs:
            pt = plan.input_pts[n]
            shape = ([self.M] + list(pt.shape[1:])) if self.plate in pt.axes else list(pt.shape)
            full[n] = torch.zeros(shape, dtype=self.dtype, device=self.device)
        for b in range(len(self.blocks)):
            lo, hi = self.blocks[b]
            tens, run = self._block(tensors, b)
            bplan = run.comp.plan
            run.dp.fwd(0, tens, lp)                                            # recompute this block's forward
            run.dp.ws_view(bplan.allreduce, self.dtype).copy_(self.tile)
            run.dp.fwd(1, tens, lp)
            outs, tmp = [], {}
            for n in bplan.grad_inputs:
                if n in bplan.global_grads:
                    tmp[n] = torch.empty(bplan.input_pts[n].shape, dtype=self.dtype, device=self.device)
                    outs.append(tmp[n])
                else:
                    outs.append(full[n][lo:hi])
            for seg in range(bplan.n_bwd):
                run.dp.bwd(seg, tens, grad_lp, outs)
            for n, g in tmp.items():
                full[n] += g
        return full


class WeightedMoments:
    """`Marginals.moments` on the device (reference Marginals.py:31-46, moments.py:16-35):
    out[plates..., *f.shape] = sum_K f(x) * w, as one factor-VM launch and one fixed-order reduction."""
    def __init__(self, x_nts: dict, w_nt: NT, f, canon, dtype, device=None):
        from . import runtime
        from .plan import weighted_moment_plan
        sizes = {}
        for v in list(x_nts.values()) + [w_nt]:
            sizes.update(v.named_sizes)
        order = lambda axes: tuple(a for a in canon if a in axes)
        self.x_axes = {k: order(v.axes) for k, v in x_nts.items()}
        self.w_axes = order(w_nt.axes)
        sigs = {k: TensorSig('sample', self.x_axes[k], v.pos_shape) for k, v in x_nts.items()}
        self.plan = weighted_moment_plan(sigs, self.w_axes, f, sizes, dtype, canon)
        self.dp = runtime.DevicePlan(self.plan, device)
        self.dtype, self.device = dtype, self.dp.device
        self.keepalive = f

    def __call__(self, x_nts: dict, w_nt: NT) -> NT:
        ins = []
        for name in self.plan.input_names:
            if name in self.plan.const_inputs:
                ins.append(self.dp.consts[name])
            elif name == '__w':
                ins.append(w_nt.order(self.w_axes).t.detach().to(self.device).to(self.dtype).contiguous())
            else:
                ins.append(x_nts[name].order(self.x_axes[name]).t.detach().to(self.device).to(self.dtype).contiguous())
        out = torch.empty(self.plan.out_shape, dtype=self.dtype, device=self.device)
        self.dp.run(0, ins, [out])
        return NT(out, self.plan.out_axes)


class StreamedRunner:
    """Host-buffer entry point: forward + backward with the outermost plate streamed through the GPU in
    `chunks` contiguous blocks, so that the host->device copy of block c+1 overlaps the kernels of block c.

    This is the reference's `Split(plate, size)` (src/alan/Split.py:44-130, logpq.py:43-57) done the B200 way:
    the blocks are the same conditionally independent shards the multi-GPU path uses (`shard_plate`), here
    laid side by side in one GPU's memory.  Each block runs forward segment 0 as soon as its inputs have
    landed; the `[K_parents]` tiles are summed (what the all-reduce does across GPUs); segment 1 and the
    backward run per block; per-element gradients are written straight into slices of the full-size
    outputs, global-parameter gradients are summed over the blocks.
    """
    def __init__(self, P, Q, sample, inputs_params, data, grad_names, stream_plate, chunks, device=None):
        self.plate, self.C = stream_plate, int(chunks)
        full = {}
        for d in (sample, inputs_params or {}, data or {}):
            full.update(d)
        M = None
        for v in full.values():
            if stream_plate in v.axes:
                M = v.named_sizes[stream_plate]
        if M is None:
            raise Exception(f"no input carries the plate {stream_plate}")
        if M % self.C:
            raise Exception(f"plate {stream_plate} of size {M} does not split into {self.C} equal blocks")
        self.M, self.m = M, M // self.C

        def head(d):
            out = {}
            for k, v in (d or {}).items():
                if stream_plate in v.axes:
                    ax = v.axes.index(stream_plate)
                    out[k] = NT(v.t.narrow(ax, 0, self.m), v.axes)
                else:
                    out[k] = v
            return out
        self.comp = Compiled(P, Q, head(sample), head(inputs_params), head(data), grad_names=list(grad_names),
                             shard_plate=stream_plate, world_size=self.C)
        plan = self.comp.plan
        if plan.n_fwd != 2:
            raise Exception(f"plate {stream_plate} is not a shardable top-level plate of this model")
        self.runs = [Runner(self.comp, device) for _ in range(self.C)]
        self.device = self.runs[0].device
        self.dtype = self.comp.dtype
        # which plan inputs carry the plate (canonical layout: plates lead, so a block is a contiguous slab)
        self.carries = []
        for name in plan.input_names:
            pt = plan.input_pts[name]
            if stream_plate in pt.axes and pt.axes[0] != stream_plate:
                raise Exception(f"{name}: the streamed plate must be the outermost axis")
            self.carries.append(stream_plate in pt.axes)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.events = [torch.cuda.Event() for _ in range(self.C)]
        self.dev = None
        self.one = torch.ones((), dtype=self.dtype, device=self.device)

    def pin(self, sample, inputs_params, data):
        """Canonical FULL-SIZE host tensors in pinned memory (what `step` consumes), in plan input order."""
        return [x.pin_memory() for x in self.comp.canonical_inputs(sample, inputs_params, data)]

    def step(self, host):
        """host: full-size canonical pinned tensors (see `pin`).  Returns (lp, {name: grad}) on the device;
        gradients of per-element parameters are full size.  The second call with the same host buffers captures
        the whole step -- the H2D copies on the copy stream, every block's programs, the tile sum -- into one CUDA
        graph and later calls replay it (no per-block host work); lp and the gradients are then static buffers."""
        if not hasattr(self, "_graphs"):
            self._graphs, self._seen, self._graph_off = {}, {}, False
        key = tuple(int(h.data_ptr()) for h in host)
        ent = self._graphs.get(key)
        if ent is not None:
            ent[0].replay()
            return ent[1], ent[2]
        if self._graph_off or self._seen.get(key, 0) < 1 or len(self._graphs) >= 2:
            if len(self._seen) > 16:
                self._seen.clear()
            self._seen[key] = self._seen.get(key, 0) + 1
            return self._step_eager(host)
        try:
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                lp, grads = self._step_eager(host)
            g.replay()
        except Exception:
            self._graph_off = True
            torch.cuda.synchronize(self.device)
            return self._step_eager(host)
        self._graphs[key] = (g, lp, grads)
        return lp, grads

    def _step_eager(self, host):
        plan = self.comp.plan
        cur = torch.cuda.current_stream(self.device)
        if self.dev is None:
            self.dev = [torch.empty_like(h, device=self.device) for h in host]
            self.grads = {n: torch.empty(([self.M] + list(plan.input_pts[n].shape[1:])) if self.plate in plan.input_pts[n].axes
                                         else list(plan.input_pts[n].shape), dtype=self.dtype, device=self.device)
                          for n in plan.grad_inputs}
        self.copy_stream.wait_stream(cur)
        with torch.cuda.stream(self.copy_stream):
            for c in range(self.C):
                for h, d, carries in zip(host, self.dev, self.carries):
                    if carries:
                        d[c * self.m:(c + 1) * self.m].copy_(h[c * self.m:(c + 1) * self.m], non_blocking=True)
                    elif c == 0:
                        d.copy_(h, non_blocking=True)
                self.events[c].record(self.copy_stream)
        tens = [[d[c * self.m:(c + 1) * self.m] if carries else d for d, carries in zip(self.dev, self.carries)]
                for c in range(self.C)]
        lp = torch.empty((), dtype=self.dtype, device=self.device)
        for c, run in enumerate(self.runs):
            cur.wait_event(self.events[c])
            run.dp.fwd(0, tens[c], lp)
        tiles = [run.dp.ws_view(plan.allreduce, self.dtype) for run in self.runs]
        total = tiles[0].clone()
        for tl in tiles[1:]:
            total += tl
        for tl in tiles:
            tl.copy_(total)
        for c, run in enumerate(self.runs):
            run.dp.fwd(1, tens[c], lp)
        gsum = {n: None for n in plan.global_grads}
        for c, run in enumerate(self.runs):
            outs = []
            for n in plan.grad_inputs:
                if n in gsum:
                    outs.append(torch.empty(plan.input_pts[n].shape, dtype=self.dtype, device=self.device))
                else:
                    outs.append(self.grads[n][c * self.m:(c + 1) * self.m])
            for seg in range(plan.n_bwd):
                run.dp.bwd(seg, tens[c], self.one, outs)
            for n, o in zip(plan.grad_inputs, outs):
                if n in gsum:
                    gsum[n] = o if gsum[n] is None else gsum[n] + o
        for n, g in gsum.items():
            self.grads[n].copy_(g)
        return lp, self.grads


class PipelinedRunner:
    """Host-batch entry point of a training loop: one log-evidence + gradient step per submitted batch of HOST
    tensors, two batches in flight.

    What the reference's loop does per iteration -- move the minibatch to the device, `elbo_rws().backward()`, read
    the loss (examples/runner.py:120-160) -- is three serial phases.  Here they run on three streams: the H2D copy
    of batch s+1 (copy stream, into the idle one of two device input sets) overlaps the kernels of batch s (compute
    stream; `Runner.step`, one replayed CUDA graph per input set) and the D2H of batch s-1's results (read-back
    stream, into that batch's pinned host buffers).  Steady state costs max(copy, compute) per step instead of their
    sum; on cfg-5 the copy (61 MB over PCIe) is the longer of the two.

        h = pipe.submit(host_tensors)        # canonical pinned host tensors (see `pin`); returns a ticket
        lp, grads = pipe.result(h)           # blocks until THAT batch's results are on the host (pinned tensors,
                                             # valid until two more batches have been submitted)
    """
    def __init__(self, comp: Compiled, device=None, process_group=None, depth: int = 2, runner: "Runner" = None):
        self.run = runner if runner is not None else Runner(comp, device, process_group)
        self.comp, self.device, self.dtype = comp, self.run.device, comp.dtype
        self.depth = int(depth)
        plan = comp.plan
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.back_stream = torch.cuda.Stream(device=self.device)
        mk = lambda: [torch.empty(plan.input_pts[n].shape, dtype=self.dtype, device=self.device) for n in plan.input_names]
        self.dev = [mk() for _ in range(self.depth)]
        self.lp_host = [torch.empty((), dtype=self.dtype).pin_memory() for _ in range(self.depth)]
        self.g_host = [{n: torch.empty(plan.input_pts[n].shape, dtype=self.dtype).pin_memory() for n in plan.grad_inputs}
                       for _ in range(self.depth)]
        self.g_dev = [None] * self.depth             # device result buffers of each slot (static once its graph exists)
        self.landed = [torch.cuda.Event() for _ in range(self.depth)]       # H2D of the slot's batch finished
        self.computed = [torch.cuda.Event() for _ in range(self.depth)]     # step on the slot finished
        self.read = [torch.cuda.Event() for _ in range(self.depth)]         # D2H of the slot's results finished
        self.free = [torch.cuda.Event() for _ in range(self.depth)]         # the slot's device results were copied out
        self.n = 0
        self.before_step = None                      # optional callable run on the compute stream before every step
        self.stage = [{} for _ in range(self.depth)]  # per slot: input index -> one-byte device staging buffer

    def pin(self, sample, inputs_params, data):
        """Canonical host tensors in pinned memory, in plan input order (what `submit` consumes).  uint8 / bool
        tensors (binary features, 0/1 observations) stay one byte per element: they cross PCIe as bytes and are
        widened on the device right after their copy (alan_b200_widen_u8)."""
        return [x.pin_memory() for x in self.comp.canonical_inputs(sample, inputs_params, data, keep_narrow=True)]

    def submit(self, host) -> int:
        from . import runtime
        slot = self.n % self.depth
        cur = torch.cuda.current_stream(self.device)
        if self.n >= self.depth:
            # the slot's previous batch: its kernels have read the device inputs, its results have left the device
            self.copy_stream.wait_event(self.computed[slot])
            cur.wait_event(self.free[slot])
            self.read[slot].synchronize()            # its host result buffers are about to be reused
        with torch.cuda.stream(self.copy_stream):
            for i, (d, h) in enumerate(zip(self.dev[slot], host)):
                if h.dtype in runtime.NARROW_DTYPES:
                    st = self.stage[slot].get(i)
                    if st is None or st.dtype != h.dtype:
                        st = self.stage[slot][i] = torch.empty(d.shape, dtype=h.dtype, device=self.device)
                    st.copy_(h, non_blocking=True)
                    runtime.widen(st, d)             # on the copy stream, straight after the bytes have landed
                else:
                    d.copy_(h, non_blocking=True)
            self.landed[slot].record(self.copy_stream)
        cur.wait_event(self.landed[slot])
        if self.before_step is not None:
            self.before_step()
        lp, grads = self.run.step(self.dev[slot])
        self.computed[slot].record(cur)
        self.back_stream.wait_event(self.computed[slot])
        with torch.cuda.stream(self.back_stream):
            self.lp_host[slot].copy_(lp, non_blocking=True)
            for n, g in grads.items():
                self.g_host[slot][n].copy_(g, non_blocking=True)
            self.free[slot].record(self.back_stream)
            self.read[slot].record(self.back_stream)
        # results of an eager (not yet captured) step are fresh tensors: keep them alive until they are copied out
        self.g_dev[slot] = (lp, grads)
        self.n += 1
        return self.n - 1

    def result(self, ticket: int):
        if ticket < self.n - self.depth or ticket >= self.n:
            raise Exception(f"batch {ticket} is not in flight (submitted so far: {self.n}, depth {self.depth})")
        slot = ticket % self.depth
        self.read[slot].synchronize()
        return self.lp_host[slot], self.g_host[slot]
