#!/usr/bin/env python
"""fan_lse on its three kernels (dense tcgen05 = fan_tc2.cuh, block-diagonal tcgen05 = fan_tc.cuh, FFMA2 =
fused.cuh) against the engine's own fp64 run on the same MovieLens-shaped inputs: relative errors of the
log-evidence and of every gradient.  Debugging / precision aid, never a bench line.
    python tools/fan_paths.py [M N K d] ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch as t
import models
from alan_b200 import model as M
from alan_b200.named import NT, from_torch_named
from alan_b200.engine import Compiled, Runner

ENV = {"tc2": {}, "tc2_f16": {"ALAN_B200_TC_F16": "1"}, "tc2_stag": {"ALAN_B200_TC_STAG": "1"},
       "tc2_f16_stag": {"ALAN_B200_TC_F16": "1", "ALAN_B200_TC_STAG": "1"},
       "tc2_nbg2": {"ALAN_B200_TC_NBG": "2"},
       "tc2_all": {"ALAN_B200_TC_F16": "1", "ALAN_B200_TC_STAG": "1", "ALAN_B200_TC_NBG": "2"}, "blockdiag": {"ALAN_B200_TC_BLOCKDIAG": "1"}, "ffma": {"ALAN_B200_NO_TC": "1"}}


def nt(d):
    return {k: from_torch_named(v) for k, v in d.items()}


def rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def case(M_, N_, K, d, seed=21):
    out = {}
    for dtype in (t.float64, t.float32):
        P, Q = models.movielens_model(M, d=d)
        inp = models.movielens_inputs(M=M_, N=N_, d=d, seed=seed, dtype=dtype)
        g = t.Generator().manual_seed(seed + 1)
        sizes = dict(inp['platesizes'])
        sample = {}
        g2p, v2g = Q.groupvarname2platenames(), Q.varname2groupvarname()
        for v, grp in v2g.items():
            axes = tuple(g2p[grp]) + (M.Kname(grp),)
            shp = [sizes[a] if a in sizes else K for a in axes] + [d]
            sample[v] = NT((0.7 * t.randn(shp, generator=g, dtype=t.float64)).to(dtype), axes)
        ip = {**nt(inp['inputs']), **nt(inp['params'])}
        data = nt(inp['data'])
        names = list(inp['params'])
        for path, env in (ENV.items() if dtype == t.float32 else [("f64", {})]):
            for k in ("ALAN_B200_TC_BLOCKDIAG", "ALAN_B200_NO_TC", "ALAN_B200_TC_F16", "ALAN_B200_TC_STAG"):
                os.environ.pop(k, None)
            os.environ.update(env)
            comp = Compiled(P, Q, sample, ip, data, grad_names=names)     # the switches are read when the plan is built
            run = Runner(comp, "cuda:0")
            tensors = run.device_inputs(sample, ip, data)
            lp = run.forward_raw(tensors)
            grads = run.backward_raw(tensors)
            t.cuda.synchronize()
            out[path] = (lp.cpu().clone(), {k: v.cpu().clone() for k, v in grads.items()})
    ref_lp, ref_g = out["f64"]
    print(f"M={M_} N={N_} K={K} d={d}: lp = {float(ref_lp):.6f}")
    for path in ENV:
        lp, g = out[path]
        print(f"  {path:10s} lp rel {rel(lp, ref_lp):.2e}   grads " +
              " ".join(f"{k}:{rel(g[k], ref_g[k]):.1e}" for k in g))
    sys.stdout.flush()


if __name__ == "__main__":
    args = [int(x) for x in sys.argv[1:]]
    cases = [args[i:i + 4] for i in range(0, len(args), 4)] or [[64, 5, 30, 18], [33, 4, 17, 18], [40, 3, 32, 18],
                                                                [50, 2, 12, 8], [300, 5, 30, 18], [2000, 10, 30, 18]]
    for c in cases:
        case(*c)
