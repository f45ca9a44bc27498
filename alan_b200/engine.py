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
from .plan import Planner, TensorSig, Plan
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


class Compiled:
    """Plan + the canonical (contiguous, canonical axis order, working dtype) input list."""
    def __init__(self, P: Plate, Q: Plate, sample, inputs_params, data, extra_log_factors=None,
                 moment_specs=(), grad_names=(), N=None, shard_plate=None, world_size=1, dtype=None):
        sample, inputs_params, data = dict(sample), dict(inputs_params or {}), dict(data or {})
        elf = dict(extra_log_factors or {})
        check_PQ(P, Q, set(data.keys()))
        self.P, self.Q = P, Q
        self.dtype = dtype or working_dtype(sample, inputs_params, data, elf)
        all_plates = P.all_platenames()
        groups = Q.groupvarnames()
        canon = list(all_plates) + [Kname(g) for g in groups]
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
        sig, self.order, self.elf_keys = {}, [], {}
        extra = []
        for key, (role, orig, v) in named.items():
            axes = tuple(a for a in canon if a in v.axes)
            sig[key] = TensorSig(role, axes, v.pos_shape, requires_grad=False)
            self.order.append((key, role, orig, axes))
            if role == 'elf':
                self.elf_keys[orig] = key
        planner = Planner(P, Q, sig, sizes, self.dtype, want_sample_N=N, shard_plate=shard_plate,
                          world_size=world_size)
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
        self.plan: Plan = planner.build(grad_names=gnames, with_sample=N is not None)
        self.planner = planner

    def _key_of(self, name):
        if name in self.elf_keys:
            return self.elf_keys[name]
        return name

    def canonical_inputs(self, sample, inputs_params, data, extra_log_factors=None, device=None):
        """Permute every tensor to canonical axis order, make it contiguous in the working dtype."""
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
                key, role, orig, axes = next(o for o in self.order if o[0] == name)
                v = elf[orig] if role == 'elf' else src[orig]
                t = v.order(axes).t
            t = t.detach()
            if device is not None:
                t = t.to(device)
            out.append(t.to(self.dtype).contiguous())
        return out
