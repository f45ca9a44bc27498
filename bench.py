#!/usr/bin/env python
"""bench.py -- fwd+bwd log-evidence throughput of the logPQ path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference algorithm's CPU port, timed on host cores

A step = one RWS iteration of the hot path on one batch of synthetic input: forward log-evidence
+ backward (gradients of every Q parameter) of the MovieLens-shaped model
(/root/reference/examples/models/movielens/movielens.py:39-74) at K=30, d=18.
Workloads (BASELINE.json configs): `cfg2` = 300 users x 5 films, `cfg5` = 10 000 users x 50 films.  The
user plate is sharded across the ranks (this replaces Split's sequential chunks): every GPU holds
10 000 users of the same model ("scaling": "weak" -- N GPUs evaluate 10 000 N users; the only
collectives are the all-reduce of the [K_mu, K_psi] tile and of the global-parameter gradients).
`--scaling strong` keeps 10 000 users in total instead.  The cfg2 numbers (the >=50x target config) are
measured in the same run at N=1 and reported under "cfg2".

Unit of work ("cell") = one (plate element, K tuple) entry of a factor tensor the reference
materialises:  W = M*K^3 (z) + M*N*K (obs) + M*K (Q of z) + K^2 + 4K.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch as t

WORKLOADS = {
    "cfg2": dict(M=300, N=5, d=18, K=30, name="movielens_300x5_d18_K30_rws_fwd_bwd"),
    "cfg5": dict(M=10000, N=50, d=18, K=30, name="movielens_10000x50_d18_K30_rws_fwd_bwd"),
}
METRIC = "fwd+bwd log-evidence evals/sec (plate-elems*K-pairs/s)"
UNIT = "cells/s"


def cells(M, N, K, **_):
    return M * K ** 3 + M * N * K + M * K + K * K + 4 * K


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p.get("bf16_tflops", 1590.0), sm_max_mhz=p.get("sm_max_mhz", 1965.0),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, sm_max_mhz=1965.0, source="fallback")


# ------------------------------------------------------------------------------------------
def make_problem(cfg, lo, hi, seed=0, dtype=t.float32):
    """Synthetic MovieLens-shaped inputs for users [lo, hi) of the workload (seeded by user index
    so that every shard count sees the same global problem)."""
    import models
    from alan_b200 import model as Mo
    from alan_b200.named import NT, from_torch_named
    M, N, d, K = cfg["M"], cfg["N"], cfg["d"], cfg["K"]
    g = t.Generator().manual_seed(seed)
    r = lambda *s: t.randn(s, generator=g, dtype=t.float32)
    glob = dict(mu_z=0.7 * r(K, d), psi_z=0.3 * r(K, d) - 0.5,
                mu_z_loc=0.1 * r(d), mu_z_ls=-0.5 + 0.1 * r(d), psi_z_loc=0.1 * r(d), psi_z_ls=-0.5 + 0.1 * r(d))
    x = (t.rand(M, N, d, generator=g) < 0.107).float()
    z = 0.7 * r(M, K, d)
    z_loc, z_ls = 0.1 * r(M, d), -0.5 + 0.1 * r(M, d)
    obs = (t.rand(M, N, generator=g) < 0.5).float()
    sl = slice(lo, hi)
    sample = {'mu_z': NT(glob['mu_z'], ('K_mu_z',)), 'psi_z': NT(glob['psi_z'], ('K_psi_z',)),
              'z': NT(z[sl].contiguous(), ('plate_1', 'K_z'))}
    ip = {'x': NT(x[sl].contiguous(), ('plate_1', 'plate_2')),
          'mu_z_loc': NT(glob['mu_z_loc'], ()), 'mu_z_ls': NT(glob['mu_z_ls'], ()),
          'psi_z_loc': NT(glob['psi_z_loc'], ()), 'psi_z_ls': NT(glob['psi_z_ls'], ()),
          'z_loc': NT(z_loc[sl].contiguous(), ('plate_1',)), 'z_ls': NT(z_ls[sl].contiguous(), ('plate_1',))}
    data = {'obs': NT(obs[sl].contiguous(), ('plate_1', 'plate_2'))}
    P, Q = models.movielens_model(Mo, d=d)
    params = ['mu_z_loc', 'mu_z_ls', 'psi_z_loc', 'psi_z_ls', 'z_loc', 'z_ls']
    return P, Q, sample, ip, data, params


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}",
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm = sorted(float(r[0]) for r in rows if r and r[0].replace('.', '').isdigit())
        if sm:
            top = sm[len(sm) // 2:]                         # samples under load = upper half
            out["sm_mhz"] = top[len(top) // 2]
            out["sm_max_mhz"] = float(rows[0][1])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, n in enumerate(names):
            if any(len(r) > 2 + i and r[2 + i].strip() == "Active" for r in rows):
                out["reasons"].append(n)
        return out


def op_model(op, itemsize):
    """Algorithmic bytes and flops of one plan op (each distinct tensor counted once)."""
    name = type(op).__name__
    seen, nbytes, flops = set(), 0, 0

    def add(pt):
        nonlocal nbytes
        if pt is not None and pt.id not in seen:
            seen.add(pt.id)
            nbytes += pt.numel * itemsize
    if name in ("ExprOp", "ExprBwdOp"):
        f = op if name == "ExprOp" else op.fwd
        pts = math.prod(d[2] for d in f.keep + f.red)
        for lf in f.codeobj.leaves:
            add(lf.pt)
        add(op.out if name == "ExprOp" else op.gleaf)
        if name == "ExprBwdOp":
            add(op.gout)
        ni = sum(1 for ins in f.codeobj.instrs if ins[0] > 1)
        flops = pts * max(ni, 1) * (5 if any(ins[0] >= 32 for ins in f.codeobj.instrs) else 1)
        if name == "ExprBwdOp":
            flops *= 3
        tag = getattr(f, "tag", "")
    elif name in ("FanLseOp", "FanLseBwdOp"):
        f = op if name == "FanLseOp" else op.fwd
        rows = math.prod(d[2] for d in f.rho) * f.kappa[2]
        for lf in (f.v, f.l, f.s):
            add(lf.pt)
        for lf, _ in f.bfactors:
            add(lf.pt)
        add(f.out)
        if name == "FanLseBwdOp":
            add(op.gout); add(op.gS)
        pts = rows * f.F
        flops = 2 * pts * f.D + 6 * pts + 2 * rows * f.D
        tag = f.tag + (":adjoint" if name == "FanLseBwdOp" else "")
    elif name == "BernDotSumOp":
        pts = math.prod(d[2] for d in op.od + op.rd)
        add(op.a.pt); add(op.b.pt); add(op.y.pt); add(op.out)
        flops, tag = pts * (2 * op.D + 12), op.tag
    elif name == "DotOp":
        pts = math.prod(d[2] for d in op.keep + op.red)
        add(op.a.pt); add(op.b.pt); add(op.out)
        flops, tag = 2 * pts, op.tag
    elif name == "NormalFanOp":
        rows = math.prod(d[2] for d in op.rows)
        for lf in (op.v, op.l, op.s):
            add(lf.pt)
        add(op.out)
        pts = rows * op.F
        flops = 2 * pts * op.D + 2 * rows * op.D
        tag = op.tag
    elif name == "ReduceOp":
        pts = math.prod(d[2] for d in op.od + op.rd)
        for lf, _ in op.factors:
            add(lf.pt)
        add(op.lse); add(op.gout); add(op.out)
        flops = pts * (len(op.factors) + (0 if op.mode == 0 else 3))
        tag = op.tag or ("adjoint" if op.mode == 3 else "sum")
    else:
        pts, tag = 0, name
    return dict(kind=name, tag=tag, bytes=nbytes, flops=flops, points=pts)


# ------------------------------------------------------------------------------------------
def run_b200(args, cfg, rank, world, local_rank, full_report=True):
    import torch.distributed as dist
    from alan_b200.engine import Compiled, Runner
    dev = t.device(f"cuda:{local_rank}")
    t.cuda.set_device(dev)
    if args.scaling == "weak" and world > 1:
        cfg = dict(cfg, M=cfg["M"] * world, name=cfg["name"] + f"_x{world}_users")
    M = cfg["M"]
    per = (M + world - 1) // world
    lo, hi = rank * per, min(M, (rank + 1) * per)
    P, Q, sample, ip, data, params = make_problem(cfg, lo, hi)
    comp = Compiled(P, Q, sample, ip, data, grad_names=params,
                    shard_plate='plate_1' if world > 1 else None, world_size=world)
    run = Runner(comp, dev)
    plan = comp.plan
    host = [x.pin_memory() for x in comp.canonical_inputs(sample, ip, data)]
    tensors = [x.to(dev) for x in host]
    flush = t.empty(256 * 1024 * 1024 // 4, dtype=t.float32, device=dev)      # > 126 MB L2
    W = cells(**cfg)

    def step():
        # Runner.step = forward_raw + backward_raw (replayed as one CUDA graph once the buffers recur);
        # ALAN_B200_EAGER_STEP=1 keeps the two eager calls
        if os.environ.get("ALAN_B200_EAGER_STEP"):
            return run.forward_raw(tensors), run.backward_raw(tensors)
        return run.step(tensors)

    def barrier():
        if world > 1:
            dist.barrier()
        t.cuda.synchronize()

    for _ in range(args.warmup):
        lp, grads = step()       # held exactly like in the timed loop: the same output buffers (and graph bindings) recur
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    sampler_t0 = time.time()
    ev = [(t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.time()
    for s, e in ev:
        flush.fill_(1.0)
        s.record()
        lp, grads = step()
        e.record()
    barrier()
    wall = time.time() - wall0
    total_ms = sum(s.elapsed_time(e) for s, e in ev)
    tt = t.tensor([total_ms], device=dev, dtype=t.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = tt.item()
    ms_per_step = total_ms / args.steps
    value = W / (ms_per_step * 1e-3)
    # the timed region lasts ~10 ms, nvidia-smi answers every ~20 ms: every rank keeps running the SAME step loop
    # (untimed, same count on all ranks: the steps contain collectives) so the sampler gets ~0.6 s under this load
    for _ in range(0 if args.device_only else min(3000, int(600.0 / max(ms_per_step, 0.05)) + 1)):
        flush.fill_(1.0)
        step()
    barrier()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["sampled"] = "nvidia-smi every 20 ms over the timed steps and an untimed continuation of the same loop (~0.6 s)"

    if args.device_only:
        print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                          "ms_per_step": ms_per_step, "config": {"workload": cfg["name"]},
                          "gpu_launches": sum(run.dp.launches[:plan.n_fwd + plan.n_bwd]) * args.steps,
                          "note": "device-only short run (profiling aid), not a bench line"}))
        sys.exit(0)
    # ---- e2e: host (pinned) buffers -> device -> fwd+bwd -> lp and gradients back to the host, through the
    # public host-buffer entry point engine.StreamedRunner (the user plate streamed in `--chunks` blocks so that
    # the H2D copy of a block overlaps the kernels of the previous one); --chunks 1 = copy everything, then run
    h2d = sum(x.numel() * x.element_size() for x in host)
    gout_host = None
    ev2 = [(t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    lp_host = t.empty((), dtype=comp.dtype).pin_memory()
    streamed = None
    if args.chunks > 1 and world == 1 and (hi - lo) % args.chunks == 0:
        from alan_b200.engine import StreamedRunner
        streamed = StreamedRunner(P, Q, sample, ip, data, params, 'plate_1', args.chunks, device=dev)
    for i in range(args.warmup + args.steps):
        k = i - args.warmup
        if k >= 0:
            flush.fill_(1.0)
            ev2[k][0].record()
        if streamed is not None:
            lp, grads = streamed.step(host)
        else:
            for dst, src in zip(tensors, host):
                dst.copy_(src, non_blocking=True)
            lp, grads = step()
        lp_host.copy_(lp, non_blocking=True)
        if gout_host is None:
            gout_host = {n: t.empty(g.shape, dtype=g.dtype).pin_memory() for n, g in grads.items()}
        for n, g in grads.items():
            gout_host[n].copy_(g, non_blocking=True)
        if k >= 0:
            ev2[k][1].record()
        t.cuda.synchronize()
    barrier()
    d2h = lp_host.element_size() + sum(g.numel() * g.element_size() for g in gout_host.values())
    e2e_ms = sum(s.elapsed_time(e) for s, e in ev2)
    tt = t.tensor([e2e_ms], device=dev, dtype=t.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_value = W / (tt.item() / args.steps * 1e-3)

    launches = sum(run.dp.launches[:plan.n_fwd + plan.n_bwd])
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "users": cfg["M"], "films": cfg["N"], "d": cfg["d"], "K": cfg["K"],
                   "cells_per_step": W, "users_per_gpu": hi - lo,
                   "parallelism": f"plate_1 sharded over {world} rank(s)",
                   "l2": "256 MB buffer written between timed steps (L2 flush)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": (f"engine.StreamedRunner.step, plate_1 in {args.chunks} blocks (H2D of block c+1 overlaps block c)"
                        if streamed is not None else "H2D copies, Runner.step (forward_raw + backward_raw as one replayed graph), D2H")},
        "gpu_launches": launches * args.steps,
        "lp": float(lp_host),
        "wall_s_timed_region": wall,
    }
    if clocks is not None:
        line["clocks"] = clocks

    # ---- roofline of the dominant kernel: per-op CUDA events in a separate profiled pass
    if rank == 0 and full_report:
        pk = peaks()
        nprog = plan.n_fwd + plan.n_bwd
        acc = {}
        reps = max(3, min(args.steps, 10))
        lp_d = t.empty((), dtype=comp.dtype, device=dev)
        one = t.ones((), dtype=comp.dtype, device=dev)
        gouts = [t.empty(plan.input_pts[n].shape, dtype=comp.dtype, device=dev) for n in plan.grad_inputs]
        for rep in range(reps + 1):
            flush.fill_(1.0)
            for prog in range(nprog):
                outs, aux = ([lp_d], []) if prog < plan.n_fwd else (gouts, [one])
                ms = run.dp.profile(prog, tensors, outs, aux)
                if rep == 0:
                    continue
                for j, m in enumerate(ms):
                    acc[(prog, j)] = acc.get((prog, j), 0.0) + m / reps
        t.cuda.synchronize()
        if acc:
            total = sum(acc.values())
            (prog, j), top_ms = max(acc.items(), key=lambda kv: kv[1])
            op = plan.programs[prog][j]
            m = op_model(op, 4)
            clk = (clocks or {}).get("sm_mhz") or pk["sm_max_mhz"]
            # pipe peaks measured on this pool's B200 (tools/microbench.cu, profiles/r01_microbench.txt):
            # 121 FMA/clk/SM (FFMA and FFMA2), 15.9 ex2/clk/SM
            fp32_peak = 148 * 121 * 2 * clk * 1e6 / 1e12                       # TFLOP/s at the clock seen under load
            mufu_peak = 148 * 15.9 * clk * 1e6                                 # ex2 / s
            t_hbm = m["bytes"] / (pk["hbm_gbs"] * 1e9)
            t_fp = m["flops"] / (fp32_peak * 1e12)
            tc_path = m["kind"] in ("FanLseOp", "FanLseBwdOp") and comp.dtype == t.float32 and \
                not os.environ.get("ALAN_B200_NO_TC")
            if tc_path:
                from alan_b200.plan import dense_fan_geometry
                fop = op if m["kind"] == "FanLseOp" else op.fwd
                dense = None if os.environ.get("ALAN_B200_TC_BLOCKDIAG") else dense_fan_geometry(fop, 4)
                roof = dict(bound="tensor", achieved=m["flops"] / (top_ms * 1e-3) / 1e12, peak=pk["bf16_tflops"],
                            unit="TFLOP/s")
                if dense is not None:
                    # fan_lse on tcgen05, dense formulation (csrc/fan_tc2.cuh): fp32-accurate 3xTF32 of the expanded
                    # square.  Algorithmic flops (2 D per cell) against the measured dense bf16 peak; what this
                    # formulation can reach is peak / 2 (tf32) / 3 (split) x D / KT (K = 2 D + 1 padded to 8) x the
                    # used share of the 128-lane tiles and 32-column user slots.
                    _, L, NG = dense
                    KT = (2 * fop.D + 1 + 7) // 8 * 8
                    tiles = -(-(L * fop.F) // 128)
                    eff = (fop.D / KT) * (L * fop.F / (128.0 * tiles)) * (fop.kappa[2] / 32.0)
                    roof["formulation_ceiling"] = pk["bf16_tflops"] / 6 * eff
                    roof["path"] = ("tcgen05.mma kind::tf32, 3xTF32 split of the expanded square, (lam, f) on TMEM lanes, "
                                    "A in TMEM (UTCHMMA / LDTM in SASS), %d fan groups" % NG)
                else:
                    # block-diagonal kernel (csrc/fan_tc.cuh): peak / 2 (tf32) / 3 (split) / 4 (block-diagonal zeros)
                    roof["formulation_ceiling"] = pk["bf16_tflops"] / 24
                    roof["path"] = "tcgen05.mma kind::tf32, 3xTF32 split, block-diagonal A in TMEM (UTCHMMA / LDTM in SASS)"
                roof["frac_of_formulation_ceiling"] = roof["achieved"] / roof["formulation_ceiling"]
            elif t_hbm >= t_fp:
                roof = dict(bound="hbm", achieved=m["bytes"] / (top_ms * 1e-3) / 1e9, peak=pk["hbm_gbs"], unit="GB/s")
            else:
                roof = dict(bound="fp32", achieved=m["flops"] / (top_ms * 1e-3) / 1e12, peak=fp32_peak, unit="TFLOP/s")
            roof["frac"] = roof["achieved"] / roof["peak"]
            # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of this command
            # at cfg-5 on one GPU (profiles/r01_fan_lse_tc2_full.md: dram__bytes_read.sum + dram__bytes_write.sum)
            ncu_traffic = {"FanLseBwdOp": 60.08e6 + 1.30e6, "FanLseOp": 24.07e6 + 0.74e6}
            traffic = ncu_traffic.get(m["kind"]) if (tc_path and cfg["M"] // max(world, 1) == 10000 and cfg["N"] == 50) else None
            roof.update(traffic=traffic, kernel=f"{m['kind']}:{m['tag']}", kernel_ms=top_ms,
                        share_of_step=top_ms / total, peak_source=pk["source"],
                        algorithmic_bytes=m["bytes"], algorithmic_flops=m["flops"],
                        hbm_frac=m["bytes"] / (top_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                        fp32_fma_frac=m["flops"] / (top_ms * 1e-3) / 1e12 / fp32_peak,
                        timing="per-op CUDA events on the launch stream, separate profiled pass, mean of %d" % reps)
            if m["kind"] in ("FanLseOp", "FanLseBwdOp"):
                roof["mufu_frac"] = m["points"] / (top_ms * 1e-3) / mufu_peak   # one ex2 per cell is the algorithmic minimum
            line["roofline"] = roof
            tops = sorted(acc.items(), key=lambda kv: -kv[1])[:6]
            line["top_ops"] = [dict(op=f"{op_model(plan.programs[p][k], 4)['kind']}:{op_model(plan.programs[p][k], 4)['tag']}",
                                    ms=round(v, 4)) for (p, k), v in tops]
    return line


def run_reference(args, cfg, sample_users=None, iters=None, warmup=None):
    """The reference algorithm's CPU port (oracle/) on the host cores.  kind="port": the reference is
    pure Python over /root/reference, which does not exist on the GPU box."""
    from oracle import logpq_oracle as O
    from alan_b200.named import NT
    t.set_num_threads(os.cpu_count() or 1)
    M = cfg["M"]
    if sample_users is None:
        sample_users = min(M, max(8, int(2.5e5 // (cfg["K"] ** 3 // 100 + cfg["N"] * 30))))
        sample_users = min(M, 300 if cfg["N"] <= 5 else 60)
    P, Q, sample, ip, data, params = make_problem(cfg, 0, sample_users)
    sub = dict(cfg, M=sample_users)
    W = cells(**sub)
    iters = iters or args.steps
    warmup = args.warmup if warmup is None else warmup

    def step():
        ipg = {k: NT(v.t.clone().requires_grad_() if k in params else v.t, v.axes) for k, v in ip.items()}
        L = O.elbo(P, Q, sample, ipg, data)
        t.autograd.grad(L, [ipg[k].t for k in params])
        return L
    for _ in range(warmup):
        step()
    t0 = time.time()
    for _ in range(iters):
        L = step()
    dt = (time.time() - t0) / iters
    return dict(value=W / dt, unit=UNIT, cores=t.get_num_threads(), kind="port",
                sample=f"{sample_users} of {M} users (one Split chunk of plate_1), {iters} fwd+bwd after {warmup} warm-up, "
                       f"{dt * 1e3:.1f} ms each", ms_per_step=dt * 1e3, lp=float(L))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=list(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--chunks", type=int, default=1,
                    help="blocks of the user plate in the e2e (host buffer) path; >1 streams them through "
                         "engine.StreamedRunner (measured slower on B200 today: per-block host work dominates)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-only", action="store_true",
                    help="skip the e2e / cfg2 / cpu legs (short command for ncu launch lists)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    cfg = WORKLOADS[args.workload]

    if args.impl == "reference":
        if rank != 0:
            return
        steps = min(args.steps, 5)
        if args.scaling == "weak" and args.gpus > 1:          # same job description as the b200 arm at this N
            cfg = dict(cfg, M=cfg["M"] * args.gpus, name=cfg["name"] + f"_x{args.gpus}_users")
        r = run_reference(args, cfg, iters=steps, warmup=min(args.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": cfg["name"], "users": cfg["M"], "films": cfg["N"], "d": cfg["d"], "K": cfg["K"]},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch.distributed as dist
    # keep stdout to the single JSON line: NCCL prints its version banner there at level VERSION, which this image
    # configures outside the environment (nccl.conf); an explicit environment value wins over the file
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    # NCCL still prints its banner on some boxes: everything this process writes to fd 1 before the JSON line goes
    # to stderr instead, so stdout carries exactly ONE line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    if not args.device_only:
        os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=t.device(f"cuda:{local_rank}"))
    line = run_b200(args, cfg, rank, world, local_rank)
    if rank == 0 and world == 1:
        if args.workload != "cfg2":
            small = run_b200(args, WORKLOADS["cfg2"], 0, 1, local_rank, full_report=True)
            line["cfg2"] = {k: small[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches", "config") if k in small}
            if "roofline" in small:
                line["cfg2"]["roofline"] = small["roofline"]
        if not args.no_cpu_baseline:
            r = run_reference(args, cfg, iters=3, warmup=1)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            if args.workload != "cfg2":
                r2 = run_reference(args, WORKLOADS["cfg2"], iters=3, warmup=1)
                line["cfg2"]["cpu_baseline"] = {k: r2[k] for k in ("value", "unit", "cores", "kind", "sample")}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)                     # teardown chatter, if any, stays off stdout too
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
