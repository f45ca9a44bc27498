#!/usr/bin/env python
"""Host-side enqueue time of one step (forward_raw / backward_raw, eager) against the device time that follows:
tells a host-bound step from a device-bound one.  Profiling aid, never a bench line.
    python tools/host_time.py [cfg5|cfg2]"""
import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch as t
import bench
from alan_b200.engine import Compiled, Runner
cfg = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg5"]
P, Q, sample, ip, data, params = bench.make_problem(cfg, 0, cfg["M"])
comp = Compiled(P, Q, sample, ip, data, grad_names=params)
run = Runner(comp, "cuda:0")
tensors = [x.cuda() for x in comp.canonical_inputs(sample, ip, data)]
for _ in range(3):
    run.forward_raw(tensors); run.backward_raw(tensors)
t.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    lp = run.forward_raw(tensors)
    t1 = time.perf_counter()
    g = run.backward_raw(tensors)
    t2 = time.perf_counter()
    t.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"host fwd enqueue {1e3*(t1-t0):.3f} ms, bwd enqueue {1e3*(t2-t1):.3f} ms, sync {1e3*(t3-t2):.3f} ms")
