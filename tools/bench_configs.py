#!/usr/bin/env python
"""Timing of BASELINE.json configs 3 and 4 (not bench lines: bench.py reports the MovieLens metric).
  cfg3  radon-shaped nested plates (States x Counties x Zips), K=10: marginals() + importance_sample(N=100)
  cfg4  Timeseries linear-Gaussian chain T=1000, K=16: elbo_nograd() + moments([ts mean, ts mean2])
GPU: alan_b200.problem API, CUDA events, median of `reps`.  CPU: the oracle port on a bounded case.
    python tools/bench_configs.py [reps]
"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch as t
import models
from alan_b200 import model as M
from alan_b200.named import NT, from_torch_named
from alan_b200.problem import Problem

reps = int(sys.argv[1]) if (__name__ == "__main__" and len(sys.argv) > 1) else 10


def gpu_time(fn, reps):
    for _ in range(3):
        fn()
    t.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); t.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def replay_time(fn, reps, inner=20):
    """Device time of `fn` alone: its launches captured once into a CUDA graph (the engine records the op DAG on
    parallel branches under capture) and replayed `inner` times back to back, so that neither Python nor launch
    latency is inside the measurement."""
    st = t.cuda.Stream()
    st.wait_stream(t.cuda.current_stream())
    with t.cuda.stream(st):
        for _ in range(3):
            fn()
    t.cuda.current_stream().wait_stream(st)
    t.cuda.synchronize()
    g = t.cuda.CUDAGraph()
    with t.cuda.graph(g, stream=st):
        fn()
    g.replay(); t.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner):
            g.replay()
        b.record(); t.cuda.synchronize()
        ts.append(a.elapsed_time(b) / inner)
    ts.sort()
    return ts[len(ts) // 2]


def radon_sample(S, C, Z, K, seed=1):
    g = t.Generator().manual_seed(seed)
    r = lambda *s: t.randn(s, generator=g)
    return {
        'global_mean': NT(0.3 * r(K), ('K_global_latents',)), 'global_log_sigma': NT(-0.3 + 0.3 * r(K), ('K_global_latents',)),
        'State_mean': NT(1.0 + 0.3 * r(S, K), ('States', 'K_State_mean')),
        'State_log_sigma': NT(-0.5 + 0.3 * r(S, K), ('States', 'K_State_log_sigma')),
        'County_mean': NT(1.0 + 0.3 * r(S, C, K), ('States', 'Counties', 'K_County_mean')),
        'County_log_sigma': NT(-0.3 + 0.2 * r(S, C, K), ('States', 'Counties', 'K_County_log_sigma')),
        'Beta_u': NT(0.3 + 0.2 * r(S, C, K), ('States', 'Counties', 'K_Beta_u')),
        'Beta_basement': NT(0.3 + 0.2 * r(S, C, K), ('States', 'Counties', 'K_Beta_basement')),
    }


def run_radon(S, C, Z, K=10, N=100):
    inp = models.radon_inputs(S=S, C=C, Z=Z)
    P, Q = models.radon_model(M)
    nt = lambda d: {k: from_torch_named(v) if any(n is not None for n in v.names) else NT(v, ()) for k, v in d.items()}
    prob = Problem(P, Q, nt(inp['data']), inputs=nt(inp['inputs']), params=nt(inp['params']), device="cuda:0")
    s = prob.sample_from(radon_sample(S, C, Z, K))
    cells = S * C * Z * K ** 4 + S * C * (4 * K + K ** 3) + S * (2 * K + K ** 3) + K
    ms_m = gpu_time(lambda: s.marginals(), reps)
    ms_i = gpu_time(lambda: s.importance_sample(N, seed=0), reps)
    # the same two calls with the inputs already resident and canonical on the device (what the API call adds is the
    # host-side canonicalisation and H2D copy of the sample, not kernel time)
    mrun = next(r for k, r in prob._runners.items() if isinstance(k, tuple) and len(k) > 4 and k[4] is None and k[2])
    mt = mrun.device_inputs(s.sample, prob.inputs_params(), prob.data,
                            {key: NT(t.zeros(mrun.comp.plan.input_pts[name].shape), mrun.comp.plan.input_pts[name].axes)
                             for key, name in mrun.comp.elf_keys.items()})
    dev_m = gpu_time(lambda: (mrun.forward_raw(mt), mrun.backward_raw(mt)), reps)
    irun = next(r for k, r in prob._runners.items() if isinstance(k, tuple) and len(k) > 4 and k[4] == N)
    it_ = irun.device_inputs(s.sample, prob.inputs_params(), prob.data)
    g = t.Generator(device="cuda:0"); g.manual_seed(0)
    us = [t.rand([irun.comp.sizes[a] for a in batch] + [N], dtype=t.float64, device="cuda:0", generator=g)
          for batch, _ in irun.comp.plan.sample_steps]
    dev_i = gpu_time(lambda: (irun.forward_raw(it_), irun.resample_raw(it_, us)), reps)
    rep_m = replay_time(lambda: (mrun.forward_raw(mt), mrun.backward_raw(mt)), reps)
    rep_i = replay_time(lambda: (irun.forward_raw(it_), irun.resample_raw(it_, us)), reps)
    return dict(config=f"cfg3 radon S={S} C={C} Z={Z} K={K}", cells=cells, marginals_ms=ms_m,
                importance_sample_ms=ms_i, marginals_device_ms=dev_m, importance_sample_device_ms=dev_i,
                marginals_graph_replay_ms=rep_m, importance_sample_graph_replay_ms=rep_i,
                launches=dict(marginals=sum(mrun.dp.launches[:mrun.comp.plan.n_fwd + mrun.comp.plan.n_bwd]),
                              importance_sample=sum(irun.dp.launches[:irun.comp.plan.n_fwd]) + irun.dp.launches[irun.comp.plan.sample_prog]),
                N=N, cells_per_s=cells / ((ms_m + ms_i) * 1e-3))


def run_timeseries(T=1000, K=16):
    inp = models.timeseries_inputs(T=T)
    P, Q = models.timeseries_model(M)
    g = t.Generator().manual_seed(2)
    sample = {'init': NT(t.randn(K, generator=g), ('K_init',)), 'ts': NT(t.randn(T, K, generator=g), ('T', 'K_ts'))}
    prob = Problem(P, Q, {'obs': from_torch_named(inp['data']['obs'])}, device="cuda:0")
    s = prob.sample_from(sample)
    moms = [('ts', models.MOMENT_FUNCS['mean']), ('ts', models.MOMENT_FUNCS['mean2'])]
    ms_e = gpu_time(lambda: s.elbo_nograd(), reps)
    ms_m = gpu_time(lambda: s.moments(moms), reps)
    cells = T * K * K + T * K + K
    erun = s._runner(())
    et = erun.device_inputs(s.sample, prob.inputs_params(), prob.data)
    dev_e = gpu_time(lambda: erun.forward_raw(et), reps)
    mrun = s._runner(moment_specs=[((v,), f) for v, f in moms])
    mt = mrun.device_inputs(s.sample, prob.inputs_params(), prob.data)
    dev_m = gpu_time(lambda: (mrun.forward_raw(mt), mrun.backward_raw(mt)), reps)
    rep_e = replay_time(lambda: erun.forward_raw(et), reps)
    rep_m = replay_time(lambda: (mrun.forward_raw(mt), mrun.backward_raw(mt)), reps)
    return dict(config=f"cfg4 timeseries T={T} K={K}", cells=cells, elbo_nograd_ms=ms_e, moments_ms=ms_m,
                elbo_nograd_device_ms=dev_e, moments_device_ms=dev_m,
                elbo_nograd_graph_replay_ms=rep_e, moments_graph_replay_ms=rep_m,
                launches=dict(elbo_nograd=sum(erun.dp.launches[:erun.comp.plan.n_fwd]),
                              moments=sum(mrun.dp.launches[:mrun.comp.plan.n_fwd + mrun.comp.plan.n_bwd])),
                cells_per_s=cells / ((ms_e + ms_m) * 1e-3))


def cpu_oracle_radon(S, C, Z, K=10):
    from oracle import logpq_oracle as O
    inp = models.radon_inputs(S=S, C=C, Z=Z)
    P, Q = models.radon_model(M)
    nt = lambda d: {k: from_torch_named(v) if any(n is not None for n in v.names) else NT(v, ()) for k, v in d.items()}
    sample = radon_sample(S, C, Z, K)
    ip = {**nt(inp['inputs']), **nt(inp['params'])}
    t0 = time.time(); O.marginals(P, Q, sample, ip, nt(inp['data'])); dt = time.time() - t0
    return dt * 1e3


def cpu_oracle_timeseries(T, K):
    from oracle import logpq_oracle as O
    inp = models.timeseries_inputs(T=T)
    P, Q = models.timeseries_model(M)
    g = t.Generator().manual_seed(2)
    sample = {'init': NT(t.randn(K, generator=g), ('K_init',)), 'ts': NT(t.randn(T, K, generator=g), ('T', 'K_ts'))}
    data = {'obs': from_torch_named(inp['data']['obs'])}
    t0 = time.time()
    O.elbo(P, Q, sample, {}, data)
    O.moments(P, Q, sample, {}, data, [(('ts',), models.MOMENT_FUNCS['mean']), (('ts',), models.MOMENT_FUNCS['mean2'])])
    return (time.time() - t0) * 1e3


if __name__ == "__main__":
    out = []
    for (S, C, Z) in ((7, 10, 10), (64, 32, 32)):
        r = run_radon(S, C, Z)
        if (S, C, Z) == (7, 10, 10):
            r["cpu_oracle_marginals_ms"] = cpu_oracle_radon(S, C, Z)
        out.append(r)
    r = run_timeseries(1000, 16)
    r["cpu_oracle_elbo_plus_moments_ms"] = cpu_oracle_timeseries(1000, 16)
    out.append(r)
    for r in out:
        print(json.dumps(r))
