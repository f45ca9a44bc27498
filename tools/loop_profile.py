#!/usr/bin/env python
"""cProfile of the mirror API's training iteration (Problem.sample -> elbo_rws -> backward -> Adam -> float(loss)) with
everything resident on the device: where the HOST time of an iteration goes.  Profiling aid, never a bench line.
    python tools/loop_profile.py [cfg2|cfg5] [iterations]"""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch as t
import bench
from alan_b200.problem import Problem
from alan_b200.named import NT

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
cfg = bench.WORKLOADS[name]
dev = t.device("cuda:0")
P, Q, sample, ip, data, params = bench.make_problem(cfg, 0, cfg["M"])
todev = lambda d: {k: NT(v.t.to(dev), v.axes) for k, v in d.items()}
par = {k: NT(v.t.clone().to(dev).requires_grad_(True), v.axes) for k, v in ip.items() if k in params}
prob = Problem(P, Q, todev(data), inputs=todev({k: v for k, v in ip.items() if k not in params}), params=par, device=dev)
opt = t.optim.Adam([v.t for v in par.values()], lr=1e-3)


def it():
    opt.zero_grad()
    L = prob.sample(cfg["K"], reparam=False).elbo_rws()
    (-L).backward()
    opt.step()
    return float(L.detach())


for _ in range(10):
    it()
t.cuda.synchronize()
import time
t0 = time.time()
for _ in range(iters):
    it()
t.cuda.synchronize()
print(f"{name}: {(time.time() - t0) / iters * 1e3:.3f} ms per iteration (wall)")
pr = cProfile.Profile()
pr.enable()
for _ in range(iters):
    it()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
