"""Adapter between live reference (alan) objects and the B200 engine.

The reference hands `logPQ_plate` a `Plate` tree of `Dist` / `Group` / `Data` / `Timeseries` objects and
trees of first-class-dim (functorch.dim) tensors (reference: src/alan/logpq.py:15-36,
src/alan/Sample.py:69-108).  This module converts both into what the engine consumes -- the
declarative tree of `alan_b200.model` and flat dicts of named tensors (`alan_b200.named.NT`) -- and
offers `B200`, the `computation_strategy` object INTEGRATION.md wires into `Sample._elbo`.

It never imports the reference: it only reads attributes of the objects it is given, so the package
imports (and the engine runs) where the reference is absent.
"""
from __future__ import annotations

import torch

from . import model as M
from .named import NT


# ---------------------------------------------------------------------------- model tree
def dist_from_reference(d) -> M.Dist:
    """reference Dist (src/alan/dist.py:102-199) -> alan_b200.model.Dist (same argument meaning)."""
    family = d.dist.__name__
    args = {}
    args.update(d.val_args)
    args.update(d.tensor_args.to_dict() if hasattr(d.tensor_args, "to_dict") else dict(d.tensor_args))
    args.update(d.str_args)
    args.update(d.func_args)
    if getattr(d, "using_sample_shape", False):
        raise Exception("sample_shape is not supported by the B200 factor kernels")
    return M.Dist(family, **args)


def _node_from_reference(v):
    if getattr(v, "is_timeseries", False):
        return M.Timeseries(v.init, dist_from_reference(v.trans))
    if hasattr(v, "dist"):
        return dist_from_reference(v)
    if type(v).__name__ == "Data":
        return M.Data()
    raise Exception(f"cannot convert {type(v)} to the B200 model tree")


def plate_from_reference(plate) -> M.Plate:
    """reference Plate (src/alan/Plate.py:50-83: grouped_prog / flat_prog) -> alan_b200.model.Plate."""
    kwargs = {}
    for k, v in plate.grouped_prog.items():
        if isinstance(v, dict):
            if len(v) >= 2:
                kwargs[k] = M.Group(**{gk: _node_from_reference(gv) for gk, gv in v.items()})
            else:
                (gk, gv), = v.items()
                kwargs[gk] = _node_from_reference(gv)
        else:
            kwargs[k] = plate_from_reference(v)
    return M.Plate(**kwargs)


# ---------------------------------------------------------------------------- tensors
def flatten_tree(tree: dict) -> dict:
    """{plate: {...}, name: tensor} -> {name: tensor} (reference Plate.flatten_tree, Plate.py:331-352)."""
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            inner = flatten_tree(v)
            dup = set(inner) & set(out)
            if dup:
                raise Exception(f"duplicate names {sorted(dup)} while flattening a tree")
            out.update(inner)
        else:
            out[k] = v
    return out


def nt_from_torchdim(x) -> NT:
    """first-class-dim tensor -> NT: named dims first (in the tensor's own dim order), positional after.
    Plain tensors come back with no named axes."""
    dims = tuple(getattr(x, "dims", ()))
    if not dims:
        return NT(x if isinstance(x, torch.Tensor) else torch.as_tensor(x), ())
    return NT(x.order(*dims), tuple(str(d) for d in dims))


def nts_from_tree(tree: dict) -> dict:
    return {k: nt_from_torchdim(v) for k, v in flatten_tree(tree).items()}


# ---------------------------------------------------------------------------- strategy object
def compile_from_reference(P, Q, sample, inputs_params, data, extra_log_factors=None, grad_names=(),
                           shard_plate=None, process_group=None, device=None, N=None):
    """Plan + workspace for one (model, shapes): returns an `engine.Runner` with `.adapter_state`."""
    from .engine import Compiled, Runner
    Pm, Qm = plate_from_reference(P), plate_from_reference(Q)
    world = 1
    if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
        world = torch.distributed.get_world_size(process_group)
    comp = Compiled(Pm, Qm, nts_from_tree(sample), nts_from_tree(inputs_params), nts_from_tree(data),
                    extra_log_factors=nts_from_tree(extra_log_factors or {}), grad_names=list(grad_names), N=N,
                    shard_plate=shard_plate if world > 1 else None, world_size=world)
    return Runner(comp, device, process_group)


def tensors_from_reference(runner, sample, inputs_params, data, extra_log_factors=None):
    """Canonical device tensors for `runner.elbo`, left on the autograd tape for `grad_names`."""
    return runner.device_inputs(nts_from_tree(sample), nts_from_tree(inputs_params), nts_from_tree(data),
                                nts_from_tree(extra_log_factors or {}), differentiable=True)


class B200:
    """`computation_strategy=B200()`: run the whole plate tree on the B200 engine.

    The reference's strategy protocol (`split_args`, src/alan/Split.py:7-14,44-71) is kept so that the
    object can be passed wherever a strategy is expected; the hook in `Sample._elbo` calls `logPQ`
    (`install_hook` below applies that two-line hook to an imported reference package).
    """
    def __init__(self, shard_plate=None, process_group=None, device=None, split=None):
        self.shard_plate, self.pg, self.device, self.split = shard_plate, process_group, device, split
        self._cache = {}

    def split_args(self, name, sample, inputs_params, extra_log_factors, data, all_platedims):
        return [dict(sample=sample, inputs_params=inputs_params, extra_log_factors=extra_log_factors,
                     data=data, all_platedims=all_platedims)]

    def logPQ(self, P, Q, sample, inputs_params, data, extra_log_factors=None, grad_names=None):
        """The reference's `logPQ_plate(name=None, ...)` (logpq.py:15-60) for a whole model: a 0-d tensor on the
        device, attached to autograd for every input that requires a gradient (Q / P parameters, reparameterised
        samples, the source terms J of marginals and moments)."""
        nts, ips, elf = nts_from_tree(sample), nts_from_tree(inputs_params), nts_from_tree(extra_log_factors or {})
        dts = nts_from_tree(data)
        if grad_names is None:
            grad_names = [k for d in (ips, nts, elf) for k, v in d.items() if v.t.requires_grad]
        shapes = lambda d: tuple(sorted(((str(k), tuple(v.t.shape), v.axes, str(v.t.dtype)) for k, v in d.items())))
        key = (id(P), id(Q), shapes(nts), shapes(ips), shapes(dts), shapes(elf), tuple(str(g) for g in grad_names))
        if key not in self._cache:
            runner = compile_from_reference(P, Q, sample, inputs_params, data, extra_log_factors,
                                            grad_names, self.shard_plate, self.pg, self.device)
            self._cache[key] = (runner, P, Q, list(elf.keys()))    # P / Q / keys kept alive: their id() is in the key
        runner = self._cache[key][0]
        return runner.elbo(tensors_from_reference(runner, sample, inputs_params, data, extra_log_factors))


def install_hook(alan_pkg):
    """Apply INTEGRATION.md's hook to an imported reference package: `Sample._elbo` dispatches to the B200
    engine when `computation_strategy` is a `B200` instance and is untouched otherwise.  Everything above it
    (`elbo_vi / elbo_rws / elbo_nograd / marginals / moments`, Sample.py:110-148,208-346) then runs on the GPU
    with no other change.  Returns a function that removes the hook again."""
    import sys
    S = sys.modules[alan_pkg.__name__ + ".Sample"].Sample
    orig = S._elbo

    def _elbo(self, sample, extra_log_factors, computation_strategy):
        if isinstance(computation_strategy, B200):
            return computation_strategy.logPQ(self.P.plate, self.Q.plate, sample, self.problem.inputs_params(),
                                              self.problem.data, extra_log_factors)
        return orig(self, sample, extra_log_factors, computation_strategy)
    S._elbo = _elbo

    def remove():
        S._elbo = orig
    return remove
