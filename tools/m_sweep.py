#!/usr/bin/env python
"""Per-op device times of the cfg-5 model at several user counts: separates each kernel's fixed cost from its
per-user cost (what strong scaling over GPUs is limited by).   python tools/m_sweep.py [M ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch as t
import bench
from alan_b200.engine import Compiled, Runner

Ms = [int(x) for x in sys.argv[1:]] or [312, 625, 1250, 2500, 5000, 10000]
rows = {}
for M in Ms:
    cfg = dict(bench.WORKLOADS["cfg5"], M=M)
    P, Q, sample, ip, data, params = bench.make_problem(cfg, 0, M)
    comp = Compiled(P, Q, sample, ip, data, grad_names=params)
    run = Runner(comp, "cuda:0")
    plan = comp.plan
    tensors = [x.cuda() for x in comp.canonical_inputs(sample, ip, data)]
    flush = t.empty(64 * 1024 * 1024, dtype=t.float32, device="cuda")
    lp_d = t.empty((), device="cuda"); one = t.ones((), device="cuda")
    gouts = [t.empty(plan.input_pts[n].shape, device="cuda") for n in plan.grad_inputs]
    acc, reps = {}, 5
    for rep in range(reps + 1):
        flush.fill_(1.0)
        for prog in range(plan.n_fwd + plan.n_bwd):
            outs, aux = ([lp_d], []) if prog < plan.n_fwd else (gouts, [one])
            ms = run.dp.profile(prog, tensors, outs, aux)
            if rep:
                for j, m in enumerate(ms):
                    acc[(prog, j)] = acc.get((prog, j), 0.0) + m / reps
    # graph-replayed whole step
    for _ in range(5):
        run.step(tensors)
    t.cuda.synchronize()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(20):
        flush.fill_(1.0); e0.record(); run.step(tensors); e1.record(); t.cuda.synchronize(); tot += e0.elapsed_time(e1)
    names = {}
    for (p, j), ms in acc.items():
        m = bench.op_model(plan.programs[p][j], 4)
        key = f"p{p}#{j:02d} {m['kind']}:{m['tag']}"[:52]
        rows.setdefault(key, {})[M] = ms * 1e3
    rows.setdefault("== whole step (graph replay) ==", {})[M] = tot / 20 * 1e3
print("%-54s" % "op  \\  users" + "".join("%9d" % M for M in Ms))
for k, v in rows.items():
    if max(v.values()) >= 8.0 or k.startswith("=="):
        print("%-54s" % k + "".join("%9.1f" % v.get(M, float('nan')) for M in Ms))
