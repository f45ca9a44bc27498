import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch as t, bench
from alan_b200.engine import Compiled, Runner
for name in ('cfg2', 'cfg5'):
    cfg = bench.WORKLOADS[name]
    P, Q, sample, ip, data, params = bench.make_problem(cfg, 0, cfg["M"])
    comp = Compiled(P, Q, sample, ip, data, grad_names=params + ['z', 'mu_z', 'psi_z'])
    run = Runner(comp, "cuda:0")
    tensors = [x.cuda() for x in comp.canonical_inputs(sample, ip, data)]
    print(name, 'ws MB', comp.plan.ws_bytes / 1e6, 'ops', [len(p) for p in comp.plan.programs])
    for _ in range(2):
        lp = run.forward_raw(tensors); g = run.backward_raw(tensors)
    t.cuda.synchronize()
    a, b = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        lp = run.forward_raw(tensors); g = run.backward_raw(tensors)
    b.record(); t.cuda.synchronize()
    print(name, 'VI fwd+bwd ms', a.elapsed_time(b) / 3, float(lp))
