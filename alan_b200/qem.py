"""In-place parameters (`OptParam`, `QEMParam`) and the QEM update  --  SURVEY.md §8 row f-4.

Mirror of what the reference's `BoundPlate` does for parameters declared as direct distribution arguments
(reference src/alan/BoundPlate.py:100-190, Param.py, dist.py:140-176) and of its QEM step
(`Sample.update_qem_params` -> `BoundPlate._update_qem_params`, Sample.py:351-355, BoundPlate.py:256-296, with the
mean <-> conventional parameter conversions of conversions.py:46-296):

  * `bind(plate, platesizes)` rewrites a model tree once: every `OptParam` / `QEMParam` argument becomes a named
    parameter `{varname}_{argname}` (or `param.name`) expanded over the plates of its variable; an OptParam's
    transformation is applied where the distribution reads it.  QEM distributions get their moving-average mean
    parameters initialised with `conv2mean` of the initial conventional parameters.
  * `QEMState.update(lr, sample)`: ONE `sample.moments(...)` call on the engine for the sufficient statistics of every
    QEM variable of that side, then one launch of `qem_update_kernel` per variable (csrc/qem.cuh, C ABI
    `alan_b200_qem_update`): moving average and conversion fused, parameters overwritten in place on the device.

Families with a conversion: Normal, Bernoulli, Poisson, Exponential, HalfNormal, Gamma, Beta (Dirichlet and
MultivariateNormal are not families of the factor VM).  Upstream `PoissonConversion.conv2mean` returns the moment
OBJECT instead of the rate (conversions.py:76-78: a bug that makes Poisson QEM fail at construction); here it is
`(rate,)`, the value `mean2conv` inverts.
"""
from __future__ import annotations

import copy

import torch

from .model import Plate, Group, Dist, Data, Timeseries, Param, OptParam, QEMParam, datagroup
from .named import NT
from . import runtime

# sufficient statistics per family (conversions.py `sufficient_stats`) and the moment functions (moments.py:81-88)
MOMENT_FUNCS = {
    'mean': lambda x: x,
    'mean2': lambda x: x * x,
    'mean_log': lambda x: x.log(),
    'mean_log1m': lambda x: (1 - x).log(),
}
SUFFICIENT = {
    'Normal': ('mean', 'mean2'), 'Bernoulli': ('mean',), 'Poisson': ('mean',), 'Exponential': ('mean',),
    'HalfNormal': ('mean2',), 'Gamma': ('mean_log', 'mean'), 'Beta': ('mean_log', 'mean_log1m'),
}
CONV_ARGS = {
    'Normal': ('loc', 'scale'), 'Bernoulli': ('probs',), 'Poisson': ('rate',), 'Exponential': ('rate',),
    'HalfNormal': ('scale',), 'Gamma': ('concentration', 'rate'), 'Beta': ('concentration1', 'concentration0'),
}


def conv2mean(family, a: dict):
    """Conventional -> mean parameters (conversions.py `conv2mean`); runs once, on the initial values."""
    if family == 'Normal':
        return (a['loc'], a['loc'] ** 2 + a['scale'] ** 2)
    if family in ('Bernoulli',):
        return (a['probs'],)
    if family == 'Poisson':
        return (a['rate'],)
    if family == 'Exponential':
        return (torch.reciprocal(a['rate']),)
    if family == 'HalfNormal':
        return (a['scale'] ** 2,)
    if family == 'Gamma':
        return (-torch.log(a['rate']) + torch.digamma(a['concentration']), a['concentration'] / a['rate'])
    if family == 'Beta':
        norm = torch.digamma(a['concentration1'] + a['concentration0'])
        return (torch.digamma(a['concentration1']) - norm, torch.digamma(a['concentration0']) - norm)
    raise Exception(f"QEM: no mean <-> conventional parameter conversion for {family}")


class QemVar:
    def __init__(self, varname, family, plates, arg2param, meannames):
        self.varname, self.family, self.plates = varname, family, tuple(plates)
        self.arg2param, self.meannames = dict(arg2param), tuple(meannames)


def expand_named(init: torch.Tensor, plates, platesizes) -> NT:
    """BoundPlate.expand_named (BoundPlate.py:17-31): the initial value broadcast over the plates of its variable."""
    for p in plates:
        if p not in platesizes:
            raise Exception(f"{p} is a plate dimension, but is not given in all_platesizes")
    shape = [platesizes[p] for p in plates]
    x = init.detach().clone()
    return NT(x.expand([*shape, *x.shape]).contiguous(), tuple(plates))


def _trans_lambda(paramname, trans):
    """A function whose ARGUMENT NAME is the parameter (how distribution lambdas name things in scope)."""
    return eval(f"lambda {paramname}: _trans({paramname})", {"_trans": trans})


def bind(plate: Plate, platesizes: dict, taken=()):
    """-> (plate with every Param replaced by a name, opt {name: NT}, qem_params {name: NT}, qem_means {name: NT},
    [QemVar])."""
    opt, qparams, qmeans, qvars = {}, {}, {}, []
    taken = set(taken)

    def new_name(name):
        if name in taken or name in opt or name in qparams:
            raise Exception(f"OptParam / QEMParam is trying to add parameter named {name}, but there's already a "
                            f"parameter with this name")
        return name

    def rebind_dist(varname, d, plates):
        if isinstance(d, Timeseries) or isinstance(d, Data) or not any(isinstance(v, Param) for v in d.args.values()):
            return d
        d2 = copy.copy(d)
        d2.args, d2.all_args = {}, []
        arg2param, conv = {}, {}
        for argname, v in d.args.items():
            if isinstance(v, Param):
                name = new_name(v.name if v.name is not None else f"{varname}_{argname}")
                val = expand_named(v.init, plates, platesizes)
                if isinstance(v, QEMParam):
                    qparams[name] = val
                    conv[argname] = val.t
                    arg2param[argname] = name
                    v = name
                else:
                    opt[name] = NT(val.t.requires_grad_(True), val.axes)
                    v = _trans_lambda(name, v.trans) if v.trans is not None else name
            d2.args[argname] = v
            if isinstance(v, str):
                d2.all_args.append(v)
            elif callable(v) and not isinstance(v, torch.Tensor):
                from .model import function_arguments
                d2.all_args.extend(function_arguments(v))
        d2.qem_dist = d2.opt_dist = False
        if d.qem_dist:
            if d.family not in SUFFICIENT:
                raise Exception(f"QEM: no mean <-> conventional parameter conversion for {d.family}")
            if set(conv) != set(CONV_ARGS[d.family]):
                raise Exception(f"QEM on {varname}: {d.family} must be parameterised by {CONV_ARGS[d.family]}")
            means = conv2mean(d.family, conv)
            names = []
            for stat, m in zip(SUFFICIENT[d.family], means):
                mn = f"{varname}_{stat}"                            # BoundPlate.py:172: f"{varname}_{moment name}"
                qmeans[mn] = NT(m.detach().clone().contiguous(), tuple(plates))
                names.append(mn)
            qvars.append(QemVar(varname, d.family, plates, arg2param, names))
        return d2

    def walk(pl: Plate, active):
        kw = {}
        for name, child in pl.grouped_prog.items():
            if isinstance(child, Plate):
                kw[name] = walk(child, (*active, name))
            elif len(child) >= 2:
                kw[name] = Group(**{k: rebind_dist(k, d, active) for k, d in child.items()})
            else:
                (k, d), = child.items()
                kw[name] = rebind_dist(k, d, active)
        return Plate(**kw)

    return walk(plate, ()), opt, qparams, qmeans, qvars


class QEMState:
    """The QEM buffers of one side (P or Q) of a Problem, resident on the device."""
    def __init__(self, qvars, qparams: dict, qmeans: dict, device, dtype):
        self.qvars = list(qvars)
        self.params = {k: NT(v.t.to(device=device, dtype=dtype).contiguous(), v.axes) for k, v in qparams.items()}
        self.means = {k: NT(v.t.to(device=device, dtype=dtype).contiguous(), v.axes) for k, v in qmeans.items()}

    def rmkeys(self):
        """[(varname, moment function)] for `sample.moments`, flat over the variables (qem_flat_list_rmkeys)."""
        return [(qv.varname, MOMENT_FUNCS[s]) for qv in self.qvars for s in SUFFICIENT[qv.family]]

    def update(self, lr: float, sample, **kw):
        if not self.qvars:
            return
        new = sample.moments(self.rmkeys(), **kw)                   # one engine call (forward + adjoint program)
        i = 0
        for qv in self.qvars:
            stats = SUFFICIENT[qv.family]
            means = [self.means[m] for m in qv.meannames]
            fresh = []
            for m in means:
                x = new[i].order(m.axes).t
                i += 1
                fresh.append(x.to(dtype=m.t.dtype, device=m.t.device).contiguous())
            params = [self.params[qv.arg2param[a]].t for a in CONV_ARGS[qv.family]]
            runtime.qem_update(qv.family, lr, fresh, [m.t for m in means], params)
