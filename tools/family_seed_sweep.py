#!/usr/bin/env python
"""Worst relative error of every density family of the factor VM (log-evidence and gradients against the oracle) over
28 seeds: shows how far tests/test_gpu_parity.py::test_density_families_vs_oracle sits from its tolerances.
    python tools/family_seed_sweep.py"""
import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch as t
import test_gpu_parity as T
from alan_b200 import model as M
from alan_b200.named import NT
from oracle import logpq_oracle as O
from golden_io import rel_err
Compiled, Runner = T._engine()
worst = {}
for family in T._FAMILIES:
  for seed in range(0, 1000, 37):
    like, gen = T._FAMILIES[family]
    dtype = t.float32
    P = M.Plate(a=M.Normal(0., 1.), b=M.Normal(-0.3, 0.5), T=M.Plate(y=like(M)))
    Q = M.Plate(a=M.Normal('a_loc', lambda a_ls: a_ls.exp()), b=M.Normal('b_loc', lambda b_ls: b_ls.exp()), T=M.Plate(y=M.Data()))
    T_, K = 23, 7
    g = t.Generator().manual_seed(seed)
    r = lambda *s: t.randn(s, generator=g, dtype=t.float64).to(dtype)
    data = {'y': NT(gen(r, T_), ('T',))}
    params = {'a_loc': NT(0.1 * r(), ()), 'a_ls': NT(-0.5 + 0.1 * r(), ()), 'b_loc': NT(-0.3 + 0.1 * r(), ()), 'b_ls': NT(-0.7 + 0.1 * r(), ())}
    sample = {'a': NT(0.6 * r(K), ('K_a',)), 'b': NT(-0.3 + 0.4 * r(K), ('K_b',))}
    names = ['a', 'b'] + list(params)
    comp = Compiled(P, Q, sample, params, data, grad_names=names)
    run = Runner(comp, "cuda:0")
    tensors = run.device_inputs(sample, params, data)
    lp = run.forward_raw(tensors); grads = run.backward_raw(tensors)
    sg = {k: NT(v.t.clone().requires_grad_(), v.axes) for k, v in sample.items()}
    pg = {k: NT(v.t.clone().requires_grad_(), v.axes) for k, v in params.items()}
    ref = O.elbo(P, Q, sg, pg, data)
    rg = t.autograd.grad(ref, [sg['a'].t, sg['b'].t] + [pg[k].t for k in params], allow_unused=True)
    e_lp = rel_err(lp.cpu(), ref)
    e_g = max([rel_err(grads[k].cpu().reshape(rr.shape), rr) for k, rr in zip(names, rg) if rr is not None])
    w = worst.get(family, (0, 0))
    worst[family] = (max(w[0], float(e_lp)), max(w[1], float(e_g)))
for f, (a, b) in worst.items():
    print(f"{f:26s} lp {a:.2e} (tol 2e-5)  grad {b:.2e} (tol 1e-3)")
