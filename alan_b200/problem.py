"""Host-side mirror of the reference's call surface for the logPQ path.

    prob = Problem(P, Q, data, inputs=..., params=...)          # reference: src/alan/Problem.py:18-69
    s = prob.sample(K) / prob.sample_from(samples)              # Problem.py:71-97 (sampling: alan_b200/sampling.py)
    s.elbo_vi() / s.elbo_rws() / s.elbo_nograd()                # Sample.py:110-148
    s.marginals(joints=...) / s.moments([...])                  # Sample.py:274-346
    marginals.moments([...]) / .ess() / .min_ess()              # Marginals.py:31-61, moments.py:16-35
    s.importance_sample(N)                                      # Sample.py:185-206
every one of them taking `computation_strategy=` (no_checkpoint | checkpoint | Split(plate, n)) like the reference.

Same names, argument meaning and error behaviour (Python `Exception` with prose) as the reference for
this path.  `sample(K)` draws from Q on the device (ancestral sampling with permuted parent particles,
SURVEY.md §8 row f-1, alan_b200/sampling.py); `sample_from` takes the reference's samples (or any
`[K, plates..., event]` tensors) instead.  Everything below `Sample` runs in the CUDA engine
through the C ABI; there is no CPU fallback.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch

from .engine import Compiled, Runner, SplitRunner, WeightedMoments
from .model import Plate, Kname, check_PQ
from .named import NT, from_torch_named
from .strategy import no_checkpoint, checkpoint, Split, resolve as _resolve_strategy
from . import runtime


def _as_nt(x) -> NT:
    if isinstance(x, NT):
        return x
    if isinstance(x, torch.Tensor):
        return from_torch_named(x) if any(n is not None for n in x.names) else NT(x, ())
    raise Exception(f"expected a tensor or NT, got {type(x)}")


class Problem:
    def __init__(self, P: Plate, Q: Plate, data: dict, inputs: Optional[dict] = None, params: Optional[dict] = None,
                 device="cuda", process_group=None, shard_plate=None, platesizes: Optional[dict] = None):
        self.P, self.Q = P, Q
        self.data = {k: _as_nt(v) for k, v in (data or {}).items()}
        self.inputs = {k: _as_nt(v) for k, v in (inputs or {}).items()}
        self.params = {k: _as_nt(v) for k, v in (params or {}).items()}
        dup = set(self.inputs) & set(self.params)
        if dup:
            raise Exception(f"names {sorted(dup)} are used both as inputs and as parameters")
        self.device = runtime.require_cuda(device)
        self.pg, self.shard_plate = process_group, shard_plate
        # compiled plans + device workspaces, shared by every Sample of this problem with the same tensor signature
        # (the reference re-walks the plate tree on every call; a plan costs ~30 ms to compile, a step ~1 ms)
        self._runners = {}
        # plate sizes: the reference takes them from BoundPlate(all_platesizes=...); here they follow from the named
        # axes of data / inputs / parameters, completed by `platesizes` for plates nothing observed spans
        self.platesizes = dict(platesizes or {})
        for d in (self.data, self.inputs, self.params):
            for k, v in d.items():
                for a, n in v.named_sizes.items():
                    if self.platesizes.setdefault(a, n) != n:
                        raise Exception(f"plate {a} has size {self.platesizes[a]} but {k} has {n} elements along it")
        # parameters declared in place (OptParam / QEMParam as distribution arguments): what the reference's BoundPlate
        # does at construction (BoundPlate.py:100-190) -- named, expanded over the plates of their variable
        from .qem import bind, QEMState
        dtype = torch.float64 if any(v.t.dtype == torch.float64 for d in (self.data, self.inputs, self.params)
                                     for v in d.values()) else torch.float32
        taken = set(self.inputs) | set(self.params)
        self.P, optP, qpP, qmP, qvP = bind(P, self.platesizes, taken)
        self.Q, optQ, qpQ, qmQ, qvQ = bind(Q, self.platesizes, taken | set(optP) | set(qpP))
        for k, v in {**optP, **optQ}.items():
            self.params[k] = NT(v.t.to(dtype).detach().requires_grad_(True), v.axes)
        self._qem = {'P': QEMState(qvP, qpP, qmP, self.device, dtype), 'Q': QEMState(qvQ, qpQ, qmQ, self.device, dtype)}
        check_PQ(self.P, self.Q, set(self.data.keys()))

    def inputs_params(self) -> dict:
        """reference Problem.inputs_params (Problem.py:113-118), flat: inputs, optimised and QEM parameters."""
        return {**self.inputs, **self.params, **self._qem['P'].params, **self._qem['Q'].params}

    def qem_params(self) -> dict:
        """Conventional parameters learned by QEM, on the device (reference BoundPlate.qem_params, BoundPlate.py:235-239)."""
        return {**self._qem['P'].params, **self._qem['Q'].params}

    def qem_means(self) -> dict:
        """Moving-average mean parameters (BoundPlate.qem_means, BoundPlate.py:241-245)."""
        return {**self._qem['P'].means, **self._qem['Q'].means}

    def update_qem_params(self, lr: float, sample, computation_strategy=None):
        """Sample.update_qem_params (Sample.py:351-355): P's QEM distributions, then Q's (whose moments already see
        P's new parameters, as upstream).  Per side: one `sample.moments` call on the engine, then the fused
        moving-average + conversion kernel per variable (alan_b200/qem.py)."""
        kw = {} if computation_strategy is None else {'computation_strategy': computation_strategy}
        with torch.no_grad():
            self._qem['P'].update(lr, sample, **kw)
            self._qem['Q'].update(lr, sample, **kw)

    def sample(self, K: int, reparam: bool = True, sampler=None, noise: Optional[dict] = None,
               seed: Optional[int] = None) -> "Sample":
        """Draw K particles per latent from Q on the device (reference Problem.sample, Problem.py:71-97 ->
        BoundPlate._sample -> Plate.sample: ancestral sampling with permuted parent particles).  One program per
        (shapes, K, sampler), one C-ABI call per draw (alan_b200/sampling.py).  `noise` / `seed` make the draw
        reproducible (explicit base noise: see QSampler.noise_shapes()).  With `reparam=True` the samples carry
        requires_grad so that `elbo_vi` returns the pathwise gradient with respect to them."""
        return Sample(self, self._draw(K, reparam, sampler, noise, seed), reparam)

    def sample_nonmp(self, K: int, reparam: bool = True, noise: Optional[dict] = None, seed: Optional[int] = None):
        """K independent draws of the whole joint from Q and the global importance-sampling estimators on them
        (reference Problem.sample_nonmp, Problem.py:99-110 -> SampleNonMP; alan_b200/nonmp.py)."""
        from .sampling import IndependentSampler
        from .nonmp import SampleNonMP
        return SampleNonMP(self, self._draw(K, reparam, IndependentSampler, noise, seed), reparam)

    def _draw(self, K, reparam, sampler, noise, seed) -> dict:
        from .sampling import QSampler, PermutationSampler
        sampler = sampler or PermutationSampler
        ip = self.inputs_params()
        dtype = torch.float64 if any(v.t.dtype == torch.float64 for d in (ip, self.data) for v in d.values()) \
            else torch.float32
        key = ('qsample', int(K), sampler, dtype, tuple(sorted((k, v.axes, tuple(v.t.shape)) for k, v in ip.items())))
        if key not in self._runners:
            self._runners[key] = QSampler(self.Q, ip, self.platesizes, K, sampler, dtype, self.device)
        qs = self._runners[key]
        with torch.no_grad():
            out = qs.run(ip, noise=noise, seed=seed)
        v2g = self.Q.varname2groupvarname()
        smp = {}
        for name, x in out.items():
            kax = Kname(v2g[name])
            axes = (kax,) + tuple(a for a in x.axes if a != kax)           # [K, plates..., event]: the reference's order
            # a VIEW in the reference's order of the program's canonical [plates, K, event] tensor: no copy here, and the
            # canonicalisation of the next elbo / marginals call finds the contiguous tensor again (two 21.6 MB copies of
            # z per iteration at cfg-5 otherwise)
            y = x.order(axes)
            smp[name] = NT(y.t.detach().requires_grad_(bool(reparam)), axes)
        return smp

    def sample_from(self, sample: dict, reparam: bool = False) -> "Sample":
        return Sample(self, {k: _as_nt(v) for k, v in sample.items()}, reparam)


class Marginals:
    """Posterior marginal weights over K (reference src/alan/Marginals.py)."""
    def __init__(self, sample: "Sample", weights: dict):
        self.sample, self.weights = sample, weights

    def ess(self) -> dict:
        """1 / sum_K w^2 per latent group (Marginals.py:48-56)."""
        out = {}
        for key, w in self.weights.items():
            if len(key) == 1:
                kdims = tuple(i for i, a in enumerate(w.axes) if a.startswith("K_"))
                out[key[0]] = NT(1.0 / (w.t * w.t).sum(kdims), tuple(a for a in w.axes if not a.startswith("K_")))
        return out

    def min_ess(self) -> float:
        """Smallest effective sample size over every marginal and plate cell (Marginals.py:58-61)."""
        ess = {}
        for key, w in self.weights.items():
            kdims = tuple(i for i, a in enumerate(w.axes) if a.startswith("K_"))
            ess[key] = 1.0 / (w.t * w.t).sum(kdims)
        return min(float(e.min()) for e in ess.values())

    def moments(self, specs):
        """specs: [(varname or tuple of varnames, f)] -> list of NT `sum_K f(x) w` with axes = plates
        (Marginals._moments_uniform_input, Marginals.py:31-46 -> RawMoment.from_marginals, moments.py:16-35).
        Every variable of one spec must belong to groups whose (joint) marginal was computed."""
        s, p = self.sample, self.sample.problem
        v2g = p.Q.varname2groupvarname()
        groups = p.Q.groupvarnames()
        canon = list(p.P.all_platenames()) + [Kname(g) for g in groups]
        out = []
        for v, f in specs:
            vs = (v,) if isinstance(v, str) else tuple(v)
            for x in vs:
                if x not in s.sample:
                    raise Exception(f"{x} is not a latent variable of Q")
            key = tuple(sorted(dict.fromkeys(v2g[x] for x in vs), key=groups.index))
            if key not in self.weights:
                raise Exception(f"the joint marginal over {key} was not computed; pass joints=[{key}] to marginals()")
            w = self.weights[key]
            xs = {x: s.sample[x] for x in vs}
            ck = ('wm', vs, id(f), tuple((k, x.axes, tuple(x.t.shape)) for k, x in xs.items()), w.axes, tuple(w.t.shape))
            cache = p._runners
            if ck not in cache:                       # the entry keeps `f` alive, so its id cannot be recycled
                cache[ck] = WeightedMoments(xs, w, f, canon, s._dtype(), p.device)
            out.append(cache[ck](xs, w))
        return out


class Sample:
    def __init__(self, problem: Problem, sample: dict, reparam: bool):
        self.problem, self.sample, self.reparam = problem, sample, reparam
        groups = problem.Q.groupvarnames()
        v2g = problem.Q.varname2groupvarname()
        for name in v2g:
            if name not in sample:
                raise Exception(f"no sample was provided for latent variable {name}")
        self.K = {}
        for name, v in sample.items():
            if name not in v2g:
                raise Exception(f"{name} is not a latent variable of Q")
            kax = Kname(v2g[name])
            if kax not in v.axes:
                raise Exception(f"sample {name} must carry its K axis {kax}")
            self.K[v2g[name]] = v.named_sizes[kax]
        self.groups = groups
        self._cache = {}

    # ------------------------------------------------------------------ engine plumbing
    def _runner(self, grad_names=(), elf=None, moment_specs=(), N=None, strategy=None):
        p = self.problem
        split = _resolve_strategy(strategy)
        sig = tuple(sorted((k, v.axes, tuple(v.t.shape), str(v.t.dtype)) for d in (self.sample, p.inputs_params(), p.data)
                           for k, v in d.items()))
        # moment functions are identified by id(): the cached runner keeps them alive (an inline lambda would
        # otherwise be freed and its address reused by the NEXT lambda, silently hitting the wrong plan)
        key = (sig, tuple(grad_names), tuple(sorted((elf or {}).keys(), key=str)),
               tuple((vs, id(f)) for vs, f in moment_specs), N,
               None if split is None else (split.platename, split.split_size))
        self._cache = p._runners
        if key not in self._cache:
            world = 1
            if p.shard_plate is not None and torch.distributed.is_available() and torch.distributed.is_initialized():
                world = torch.distributed.get_world_size(p.pg)
            if split is not None and N is None:
                run = SplitRunner(p.P, p.Q, self.sample, p.inputs_params(), p.data, split.platename, split,
                                  extra_log_factors=elf, moment_specs=moment_specs, grad_names=list(grad_names),
                                  device=p.device)
            else:
                # resampling keeps every factor of the tree in one workspace: Split is accepted and runs unsplit
                comp = Compiled(p.P, p.Q, self.sample, p.inputs_params(), p.data, extra_log_factors=elf,
                                moment_specs=moment_specs, grad_names=list(grad_names), N=N,
                                shard_plate=p.shard_plate if world > 1 else None, world_size=world)
                run = Runner(comp, p.device, p.pg)
            run._keepalive = [f for _, f in moment_specs]
            self._cache[key] = run
        return self._cache[key]

    def _elbo(self, grad_names, strategy=None):
        run = self._runner(grad_names, strategy=strategy)
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
    def elbo_vi(self, computation_strategy=checkpoint):
        """Reparameterised ELBO: gradients flow to parameters and to the samples (Sample.py:110-122)."""
        if not self.reparam:
            raise Exception("To compute the ELBO with the right gradients for VI you must construct a "
                            "reparameterised sample using `problem.sample(K, reparam=True)`")
        return self._elbo(self._diff_names(True), computation_strategy)

    def elbo_rws(self, computation_strategy=checkpoint):
        """Samples detached; gradients flow to the parameters only (Sample.py:124-134)."""
        return self._elbo(self._diff_names(False), computation_strategy)

    def elbo_nograd(self, computation_strategy=checkpoint):
        """No gradients at all (Sample.py:136-148); Split still bounds the workspace."""
        with torch.no_grad():
            return self._elbo((), computation_strategy)

    def _J_axes(self, key):
        g2p = self.problem.Q.groupvarname2platenames()
        gs = tuple(sorted(key, key=self.groups.index))
        plates = g2p[gs[0]]
        for g in gs[1:]:
            if tuple(g2p[g]) != tuple(plates):
                raise Exception(f"joint marginal {key}: groups must live in the same plates")
        return tuple(Kname(g) for g in gs) + tuple(plates)

    def marginals(self, joints: Sequence = (), computation_strategy=checkpoint) -> Marginals:
        """Posterior marginals over K for every latent group (+ the requested joints): the gradient of the
        log-evidence w.r.t. zero source terms J (Sample.py:208-289)."""
        v2g = self.problem.Q.varname2groupvarname()
        keys = [(g,) for g in self.groups]
        for j in joints:
            gs = tuple(dict.fromkeys(v2g.get(x, x) for x in j))
            for g in gs:
                if g not in self.groups:
                    raise Exception(f"{g} is not a latent variable or group of Q")
            keys.append(tuple(sorted(gs, key=self.groups.index)))
        sizes = self._sizes()
        dtype = self._dtype()
        elf = {}
        for key in keys:
            axes = self._J_axes(key)
            elf[key] = NT(torch.zeros([sizes[a] for a in axes], dtype=dtype), axes)
        run = self._runner(grad_names=list(elf.keys()), elf=elf, strategy=computation_strategy)
        p = self.problem
        tens = run.device_inputs(self.sample, p.inputs_params(), p.data, elf)
        run.forward_raw(tens)
        grads = run.backward_raw(tens)
        out = {}
        for key in keys:
            name = run.comp.elf_keys[key]
            out[key] = NT(grads[name], run.comp.plan.input_pts[name].axes)
        return Marginals(self, out)

    def moments(self, specs, computation_strategy=no_checkpoint):
        """specs: [(varname or tuple of varnames, f)] -> list of NT `E_post[f(x)]` with axes = plates of the
        variables (Sample.py:291-346; gradient w.r.t. the zero source term of the factor sum f(x) * J)."""
        moms = [((v,) if isinstance(v, str) else tuple(v), f) for v, f in specs]
        for vs, _ in moms:
            for v in vs:
                if v not in self.sample:
                    raise Exception(f"{v} is not a latent variable of Q")
        run = self._runner(moment_specs=moms, strategy=computation_strategy)
        p = self.problem
        tens = run.device_inputs(self.sample, p.inputs_params(), p.data)
        run.forward_raw(tens)
        grads = run.backward_raw(tens)
        return [NT(grads[j], plates) for j, plates, _ in run.comp.moment_inputs]

    def update_qem_params(self, lr: float, computation_strategy=no_checkpoint):
        """reference Sample.update_qem_params (Sample.py:351-355)."""
        self.problem.update_qem_params(lr, self, computation_strategy)

    def importance_sample(self, N: int, uniforms=None, seed: Optional[int] = None, computation_strategy=checkpoint) -> dict:
        """N joint posterior samples: K indices drawn top-down over the plate tree, then gathered
        (Sample.py:150-206).  `uniforms` (one float64 tensor `[plates..., N]` per sampling step, in
        `plan.sample_steps` order) makes the draw reproducible and bit-comparable; by default they
        are drawn on the device from `seed`."""
        if N < 1:
            raise Exception("importance_sample needs N >= 1")
        run = self._runner(N=N, strategy=computation_strategy)
        p = self.problem
        plan = run.comp.plan
        tens = run.device_inputs(self.sample, p.inputs_params(), p.data)
        run.forward_raw(tens)
        if uniforms is None:
            g = torch.Generator(device=run.device)
            g.manual_seed(0 if seed is None else seed)
            uniforms = [torch.rand([run.comp.sizes[a] for a in batch] + [N], dtype=torch.float64, device=run.device,
                                   generator=g) for batch, _ in plan.sample_steps]
        idx = run.resample_raw(tens, uniforms)
        self.indices = idx
        v2g = p.Q.varname2groupvarname()
        g2p = p.Q.groupvarname2platenames()
        out = {}
        for name, x in self.sample.items():
            grp = v2g[name]
            plates = tuple(g2p[grp])
            xc = x.order(plates + (Kname(grp),)).t.to(run.device).contiguous()
            outer = math.prod(run.comp.sizes[a] for a in plates)
            K = xc.shape[len(plates)]
            inner = xc.numel() // max(outer * K, 1)
            got = runtime.gather(xc, idx[grp].t.reshape(N, outer), outer, K, inner)
            out[name] = NT(got.reshape([N] + [run.comp.sizes[a] for a in plates] + list(x.pos_shape)),
                           ('N',) + plates)
        from .predict import ImportanceSample
        return ImportanceSample(p, out, N)             # a dict {varname: NT} with .extend() / .moments() / .dump()

    # ------------------------------------------------------------------ helpers
    def _sizes(self):
        sizes = {}
        p = self.problem
        for d in (self.sample, p.inputs_params(), p.data):
            for v in d.values():
                sizes.update(v.named_sizes)
        return sizes

    def _dtype(self):
        p = self.problem
        for d in (self.sample, p.inputs_params(), p.data):
            for v in d.values():
                if v.t.dtype == torch.float64:
                    return torch.float64
        return torch.float32
