"""Plate sharding across ranks, world_size 2 over gloo on CPU.

Each rank compiles the plan for ITS shard of the outermost plate, runs the forward segment up
to the plate sum, all-reduces the per-shard tile (the one exchange step that replaces Split's
sequential chunk loop, reference logpq.py:151-153), finishes the top-level contraction
redundantly, and back-propagates; gradients of global (unsharded) tensors are all-reduced once.
Kernel semantics come from tests/plan_emulator.py (no GPU here); the collective is real."""
import os
import sys

import pytest
import torch as t
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret, fused=False):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import models
    from alan_b200 import model as M
    from alan_b200.engine import Compiled
    from alan_b200.named import NT, from_torch_named
    from plan_emulator import Emu

    t.manual_seed(0)
    dtype = t.float64
    Mu, N, d, K = 11, 3, 6, 5                       # 11 users: ragged shards (6 + 5)
    P, Q = models.movielens_model(M, d=d)
    inp = models.movielens_inputs(M=Mu, N=N, d=d, seed=2, dtype=dtype)
    g = t.Generator().manual_seed(5)
    r = lambda *s: (0.7 * t.randn(s, generator=g, dtype=t.float64)).to(dtype)
    full_sample = {'mu_z': NT(r(K, d), ('K_mu_z',)), 'psi_z': NT(r(K, d) - 0.5, ('K_psi_z',)),
                   'z': NT(r(Mu, K, d), ('plate_1', 'K_z'))}
    full_ip = {k: from_torch_named(v) for k, v in {**inp['inputs'], **inp['params']}.items()}
    full_data = {k: from_torch_named(v) for k, v in inp['data'].items()}
    per = (Mu + world - 1) // world
    lo, hi = rank * per, min(Mu, (rank + 1) * per)

    def shard(d_):
        out = {}
        for k, v in d_.items():
            if 'plate_1' in v.axes:
                i = v.axes.index('plate_1')
                out[k] = NT(v.t.narrow(i, lo, hi - lo).contiguous(), v.axes)
            else:
                out[k] = v
        return out
    sample, ip, data = shard(full_sample), shard(full_ip), shard(full_data)
    names = list(inp['params']) + ['mu_z', 'psi_z', 'z']           # VI-style: global AND sharded grads
    comp = Compiled(P, Q, sample, ip, data, grad_names=names, shard_plate='plate_1', world_size=world,
                    fused_collectives=fused)
    plan = comp.plan
    assert plan.allreduce is not None
    inputs = comp.canonical_inputs(sample, ip, data)
    lp = t.zeros(1, dtype=dtype)
    emu = Emu(plan, inputs, outputs={0: lp})
    if fused:
        # the exchanges are ops inside the programs (plan.XReduceOp; the emulator performs them over gloo):
        # one forward and one backward program, like an unsharded plan
        # (small reductions next to them ride in the same single-CTA launch: plan.ReduceSeqOp -- flattened here)
        flat = lambda prog: [m for op in prog for m in (op.ops if type(op).__name__ == 'ReduceSeqOp' else [op])]
        kinds = [[type(op).__name__ for op in flat(prog)] for prog in plan.programs]
        assert plan.n_fwd == 1 and plan.n_bwd == 1 and 'XReduceOp' in kinds[0] and kinds[1][-1] == 'XReduceOp'
        emu.run(plan.programs[0])
    else:
        assert plan.n_fwd == 2
        emu.run(plan.programs[0])
        tile_pt = plan.allreduce
        tile = emu.ws[tile_pt.offset // 8: tile_pt.offset // 8 + tile_pt.numel]
        dist.all_reduce(tile)                                        # the single forward exchange
        emu.run(plan.programs[1])
    emu.outputs = {i: t.zeros(plan.input_pts[n].numel, dtype=dtype) for i, n in enumerate(plan.grad_inputs)}
    emu.aux = {0: t.ones(1, dtype=dtype)}
    for seg in plan.programs[plan.n_fwd:plan.n_fwd + plan.n_bwd]:
        emu.run(seg)
    grads = {n: emu.outputs[i].reshape(plan.input_pts[n].shape) for i, n in enumerate(plan.grad_inputs)}
    assert set(plan.global_grads) == {n for n in names if 'plate_1' not in plan.input_pts[n].axes}
    if not fused:
        for n in plan.global_grads:                                  # the single backward exchange
            dist.all_reduce(grads[n])
    if rank == 0:
        from oracle import logpq_oracle as O
        sg = {k: NT(v.t.clone().requires_grad_(), v.axes) for k, v in full_sample.items()}
        pg = {k: NT(v.t.clone().requires_grad_() if k in names else v.t, v.axes) for k, v in full_ip.items()}
        ref = O.elbo(P, Q, sg, pg, full_data)
        leaves = {**{k: sg[k] for k in ('mu_z', 'psi_z', 'z')}, **{k: pg[k] for k in inp['params']}}
        rg = dict(zip(leaves, t.autograd.grad(ref, [v.t for v in leaves.values()])))
        err = {'lp': abs((lp[0] - ref).item()) / abs(ref.item())}
        for n in names:
            pt = plan.input_pts[n]
            mine = NT(grads[n], pt.axes).order(leaves[n].axes).t if pt.axes else grads[n]
            want = rg[n]
            if 'plate_1' in leaves[n].axes:
                want = want.narrow(leaves[n].axes.index('plate_1'), lo, hi - lo)
            err[n] = ((mine - want).abs().max() / want.abs().max().clamp(min=1e-300)).item()
        ret.put(err)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fused", [False, True])
def test_plate_sharding_two_ranks_gloo(fused):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() + 7 * int(fused)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret, fused)) for r in range(2)]
    for p in procs:
        p.start()
    err = ret.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k, e in err.items():
        assert e < 1e-10, (k, e)
