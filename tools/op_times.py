#!/usr/bin/env python
"""Per-op device times of one workload (CUDA events around every op via alan_b200_profile).
Profiling aid: prints a table, never a bench line.   python tools/op_times.py [cfg5|cfg2] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch as t
import bench
from alan_b200.engine import Compiled, Runner

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
if name.startswith("radon"):
    # cfg3: marginals of the radon-shaped model (grad w.r.t. zero source terms), default or scaled shape
    import models, bench_configs as BC
    from alan_b200 import model as M
    from alan_b200.named import NT, from_torch_named
    S, C, Z = (64, 32, 32) if name == "radon_big" else (7, 10, 10)
    inp = models.radon_inputs(S=S, C=C, Z=Z)
    P, Q = models.radon_model(M)
    nt = lambda d: {k: from_torch_named(v) if any(n is not None for n in v.names) else NT(v, ()) for k, v in d.items()}
    sample, ip, data = BC.radon_sample(S, C, Z, 10), {**nt(inp['inputs']), **nt(inp['params'])}, nt(inp['data'])
    g2p, sizes = Q.groupvarname2platenames(), {**inp['platesizes']}
    for v in sample.values():
        sizes.update(v.named_sizes)
    elf = {}
    for grp in Q.groupvarnames():
        axes = ('K_' + grp,) + tuple(g2p[grp])
        elf[(grp,)] = NT(t.zeros([sizes[a] for a in axes]), axes)
    comp = Compiled(P, Q, sample, ip, data, extra_log_factors=elf, grad_names=list(elf.keys()))
    run = Runner(comp, "cuda:0")
    plan = comp.plan
    tensors = [x.cuda() for x in comp.canonical_inputs(sample, ip, data, elf)]
elif name.startswith("ts"):
    # cfg4: Timeseries T=1000, K=16; "ts" = elbo (forward only), "tsm" = moments (forward + adjoint)
    import models
    from alan_b200 import model as M
    from alan_b200.named import NT, from_torch_named
    T_, K_ = 1000, 16
    inp = models.timeseries_inputs(T=T_)
    P, Q = models.timeseries_model(M)
    g = t.Generator().manual_seed(2)
    sample = {'init': NT(t.randn(K_, generator=g), ('K_init',)), 'ts': NT(t.randn(T_, K_, generator=g), ('T', 'K_ts'))}
    data = {'obs': from_torch_named(inp['data']['obs'])}
    moms = [(('ts',), models.MOMENT_FUNCS['mean']), (('ts',), models.MOMENT_FUNCS['mean2'])] if name == "tsm" else ()
    comp = Compiled(P, Q, sample, {}, data, moment_specs=moms)
    run = Runner(comp, "cuda:0")
    plan = comp.plan
    tensors = run.device_inputs(sample, {}, data)
else:
    vi = name.endswith("vi")                  # e.g. cfg5vi: gradients w.r.t. the samples too (reparameterised path)
    cfg = bench.WORKLOADS[name[:-2] if vi else name]
    P, Q, sample, ip, data, params = bench.make_problem(cfg, 0, cfg["M"])
    comp = Compiled(P, Q, sample, ip, data, grad_names=params + (['z', 'mu_z', 'psi_z'] if vi else []))
    run = Runner(comp, "cuda:0")
    plan = comp.plan
    tensors = [x.cuda() for x in comp.canonical_inputs(sample, ip, data)]
flush = t.empty(64 * 1024 * 1024, dtype=t.float32, device="cuda")
lp_d = t.empty((), dtype=comp.dtype, device="cuda")
one = t.ones((), dtype=comp.dtype, device="cuda")
gouts = [t.empty(plan.input_pts[n].shape, dtype=comp.dtype, device="cuda") for n in plan.grad_inputs]
acc = {}
for rep in range(reps + 1):
    flush.fill_(1.0)
    for prog in range(plan.n_fwd + plan.n_bwd):
        outs, aux = ([lp_d], []) if prog < plan.n_fwd else (gouts, [one])
        ms = run.dp.profile(prog, tensors, outs, aux)
        if rep:
            for j, m in enumerate(ms):
                acc[(prog, j)] = acc.get((prog, j), 0.0) + m / reps
tot = sum(acc.values())
print(f"{name}: sum of per-op times {tot:.3f} ms over {len(acc)} ops")
for (p, j), ms in acc.items():
    m = bench.op_model(plan.programs[p][j], 4)
    gbs = m["bytes"] / (ms * 1e-3) / 1e9 if ms > 0 else 0
    print(f"  p{p} #{j:02d} {ms*1e3:9.1f} us {100*ms/tot:5.1f}%  {m['kind']:12s} {m['tag']:30s} pts={m['points']:>10} "
          f"alg_bytes={m['bytes']:>10} ({gbs:7.1f} GB/s)")
