"""Global importance sampling on the same kernels: the mirror of the reference's `SampleNonMP`
(reference src/alan/SampleNonMP.py:13-125) -- the papers' comparison baseline.

    s = problem.sample_nonmp(K)                       # Problem.py:99-110: K independent draws of the whole joint
    s.elbo_vi() / s.elbo_rws() / s.elbo_nograd()      # logsumexp_K(log P - log Q) - log K      (SampleNonMP.py:56-69)
    s.moments([...])                                  # sum_K softmax(lpq)_k f(x_k)            (:100-116)
    s.importance_sample(N)                            # N categorical draws over K, gathered   (:71-98)
    s.update_qem_params(lr)                           # :121-125 (alan_b200/qem.py)

Every latent carries ONE shared K axis (`unify_dims`, SampleNonMP.py:127-137); nothing is contracted, so the plan is
the factor kernels, the plate sums and one logsumexp over K (`Planner.plan_nonmp`).  Moments are the gradient of that
logsumexp with respect to zero source terms `f(x) * J`, like the massively parallel path (the softmax weights are
exactly the reference's `(lpq - lpq.logsumexp(K)).exp()`).  No CPU fallback.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from .engine import Compiled, Runner
from .model import Kname
from .named import NT
from .plan import NONMP_K
from . import runtime


def unify_K(sample: dict) -> dict:
    """Rename the K axis of every sample tensor to the one shared axis (reference `unify_dims`)."""
    out, K = {}, None
    for k, v in sample.items():
        ks = [a for a in v.axes if a.startswith('K_')]
        if len(ks) != 1:
            raise Exception(f"sample {k} must carry exactly one K axis, has {ks}")
        if K is None:
            K = v.named_sizes[ks[0]]
        if v.named_sizes[ks[0]] != K:
            raise Exception(f"SampleNonMP needs the same K for every latent; {k} has {v.named_sizes[ks[0]]}, others {K}")
        out[k] = NT(v.t, tuple(NONMP_K if a == ks[0] else a for a in v.axes))
    return out


class SampleNonMP:
    def __init__(self, problem, sample: dict, reparam: bool):
        self.problem, self.reparam = problem, reparam
        v2g = problem.Q.varname2groupvarname()
        for name in v2g:
            if name not in sample:
                raise Exception(f"no sample was provided for latent variable {name}")
        for name in sample:
            if name not in v2g:
                raise Exception(f"{name} is not a latent variable of Q")
        self.sample = unify_K(sample)
        self.K = next(iter(self.sample.values())).named_sizes[NONMP_K]

    # ------------------------------------------------------------------ engine plumbing
    def _runner(self, grad_names=(), moment_specs=(), N=None):
        p = self.problem
        sig = tuple(sorted((k, v.axes, tuple(v.t.shape), str(v.t.dtype)) for d in (self.sample, p.inputs_params(), p.data)
                           for k, v in d.items()))
        key = ('nonmp', sig, tuple(grad_names), tuple((vs, id(f)) for vs, f in moment_specs), N)
        cache = p._runners
        if key not in cache:
            comp = Compiled(p.P, p.Q, self.sample, p.inputs_params(), p.data, moment_specs=moment_specs,
                            grad_names=list(grad_names), N=N, nonmp=True)
            run = Runner(comp, p.device, None)
            run._keepalive = [f for _, f in moment_specs]          # id(f) stays unique while the entry lives
            cache[key] = run
        return cache[key]

    def _elbo(self, grad_names):
        run = self._runner(grad_names)
        p = self.problem
        tens = run.device_inputs(self.sample, p.inputs_params(), p.data, differentiable=bool(grad_names))
        if not grad_names:
            return run.forward_raw(tens)
        return run.elbo(tens)

    def _diff_names(self, with_sample):
        names = [k for k, v in self.problem.params.items() if v.t.requires_grad]
        if with_sample:
            names += [k for k, v in self.sample.items() if v.t.requires_grad]
        return names

    # ------------------------------------------------------------------ reference surface
    def elbo_vi(self):
        if not self.reparam:
            raise Exception("To compute the ELBO with the right gradients for VI you must construct a "
                            "reparameterised sample using `problem.sample(K, reparam=True)`")
        return self._elbo(self._diff_names(True))

    def elbo_rws(self):
        return self._elbo(self._diff_names(False))

    def elbo_nograd(self):
        with torch.no_grad():
            return self._elbo(())

    def moments(self, specs, computation_strategy=None):
        """specs: [(varname or tuple of varnames, f)] -> list of NT `sum_K w_k f(x_k)` with axes = the plates of the
        variables, w = softmax over K of log P - log Q (SampleNonMP.py:100-116)."""
        moms = [((v,) if isinstance(v, str) else tuple(v), f) for v, f in specs]
        for vs, _ in moms:
            for v in vs:
                if v not in self.sample:
                    raise Exception(f"{v} is not a latent variable of Q")
        run = self._runner(moment_specs=moms)
        p = self.problem
        tens = run.device_inputs(self.sample, p.inputs_params(), p.data)
        run.forward_raw(tens)
        grads = run.backward_raw(tens)
        return [NT(grads[j], plates) for j, plates, _ in run.comp.moment_inputs]

    def importance_sample(self, N: int, uniforms=None, seed: Optional[int] = None):
        """N joint posterior samples: indices over the one K axis drawn by inverse CDF from explicit float64
        uniforms `[N]` (drawn on the device from `seed` by default), every latent gathered at them."""
        if N < 1:
            raise Exception("importance_sample needs N >= 1")
        run = self._runner(N=N)
        p = self.problem
        tens = run.device_inputs(self.sample, p.inputs_params(), p.data)
        run.forward_raw(tens)
        if uniforms is None:
            g = torch.Generator(device=run.device)
            g.manual_seed(0 if seed is None else seed)
            uniforms = [torch.rand([N], dtype=torch.float64, device=run.device, generator=g)]
        elif isinstance(uniforms, torch.Tensor):
            uniforms = [uniforms]
        idx = run.resample_raw(tens, uniforms)[NONMP_K]
        self.indices = idx
        out = {}
        for name, x in self.sample.items():
            plates = tuple(a for a in x.axes if a != NONMP_K)
            xc = x.order((NONMP_K,) + plates).t.detach().to(run.device).contiguous()
            inner = xc.numel() // max(self.K, 1)
            got = runtime.gather(xc, idx.t.reshape(N, 1), 1, self.K, inner)
            out[name] = NT(got.reshape([N] + list(xc.shape[1:])), ('N',) + plates)
        from .predict import ImportanceSample
        return ImportanceSample(p, out, N)

    def update_qem_params(self, lr: float):
        """Reference SampleNonMP.update_qem_params (SampleNonMP.py:121-125)."""
        self.problem.update_qem_params(lr, self)
