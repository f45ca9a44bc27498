"""Declarative model tree consumed by the B200 plan compiler.

This is the host-side mirror of the reference's model-declaration interface for
the logPQ path: same names, same argument meaning (reference:
src/alan/Plate.py:16-83, Group.py, Data.py, Timeseries.py:17-87, dist.py:76-206).
It is purely structural -- it holds no tensors that change per call, does no
sampling and no arithmetic.  `alan_b200.alan_adapter` converts a live reference
``Plate`` into this form; tests and the benchmark build it directly so that they
run where the reference is absent (the GPU box).

Distribution arguments follow dist.py:76-92: a number, a tensor constant, a
string naming something in scope, or a Python function whose ARGUMENT NAMES name
things in scope (``lambda psi_z: psi_z.exp()``).
"""
from __future__ import annotations

import inspect
import numbers
import types
from typing import Optional

import torch

# family -> ordered argument names, as torch.distributions binds positionals
# (reference dist.py:110 uses inspect.signature(self.dist).bind).
FAMILIES = {
    "Normal": ("loc", "scale"),
    "Bernoulli": ("probs", "logits"),
    "LogNormal": ("loc", "scale"),
    "Laplace": ("loc", "scale"),
    "Exponential": ("rate",),
    "Gamma": ("concentration", "rate"),
    "Beta": ("concentration1", "concentration0"),
    "Poisson": ("rate",),
    "Cauchy": ("loc", "scale"),
    "HalfNormal": ("scale",),
    "Uniform": ("low", "high"),
    "StudentT": ("df", "loc", "scale"),
    "NegativeBinomial": ("total_count", "probs", "logits"),
    "Binomial": ("total_count", "probs", "logits"),
    "MultivariateNormal": ("loc", "covariance_matrix", "precision_matrix", "scale_tril"),
    "Dirichlet": ("concentration",),
    # densities composed from the factor VM's primitive operations (plan.py COMPOSED)
    "Gumbel": ("loc", "scale"),
    "Weibull": ("scale", "concentration"),
    "Pareto": ("scale", "alpha"),
    "HalfCauchy": ("scale",),
    "Chi2": ("df",),
    "Geometric": ("probs", "logits"),
    "Kumaraswamy": ("concentration1", "concentration0"),
    "FisherSnedecor": ("df1", "df2"),
    "RelaxedBernoulli": ("temperature", "probs", "logits"),
    "OneHotCategorical": ("probs", "logits"),
    "Categorical": ("probs", "logits"),
    "ContinuousBernoulli": ("probs", "logits"),
    "VonMises": ("loc", "concentration"),
    "LowRankMultivariateNormal": ("loc", "cov_factor", "cov_diag"),
    "RelaxedOneHotCategorical": ("temperature", "probs", "logits"),
    "Multinomial": ("total_count", "probs", "logits"),
}
# arguments that torch.distributions leaves as None unless given
_OPTIONAL = {"probs", "logits"}
_MATRIX_ARGS = ("covariance_matrix", "precision_matrix", "scale_tril")
# arguments whose constraint is discrete (dist.py:311-318 keeps these as ints)
DISCRETE_ARGS = {("NegativeBinomial", "total_count"), ("Binomial", "total_count")}


def function_arguments(f):
    """Names of the positional arguments of `f` (reference utils.function_arguments)."""
    return tuple(inspect.signature(f).parameters.keys())


class Param:
    """A parameter declared in place, as a direct argument of a distribution (reference src/alan/Param.py)."""
    def __init__(self, init, ignore_platenames=(), name=None):
        if isinstance(init, numbers.Number):
            init = torch.tensor(float(init))
        if not isinstance(init, torch.Tensor):
            raise Exception("the initial value of an OptParam / QEMParam must be a number or a tensor")
        self.init, self.ignore_platenames, self.name = init.detach(), tuple(ignore_platenames), name
        self.trans = None


class OptParam(Param):
    """Learned by optimisation: `Normal(OptParam(0.), OptParam(1., transformation=t.exp))` (Param.py:17-25)."""
    def __init__(self, init, transformation=None, ignore_platenames=(), name=None):
        super().__init__(init, ignore_platenames, name)
        self.trans = transformation


class QEMParam(Param):
    """Learned by QEM, the moving average of posterior moments (Param.py:27-32, BoundPlate.py:256-296)."""


class Dist:
    """One distribution node: a family plus unresolved arguments."""
    is_timeseries = False

    def __init__(self, family: str, *args, **kwargs):
        if family not in FAMILIES:
            raise Exception(f"distribution family {family} is not supported by the B200 factor kernel")
        names = FAMILIES[family]
        bound = {}
        if len(args) > len(names):
            raise Exception(f"Wrong number of arguments provided to {family}")
        for n, a in zip(names, args):
            bound[n] = a
        for k, v in kwargs.items():
            if k not in names:
                raise Exception(f"{family} has no argument {k}")
            if k in bound:
                raise Exception(f"{family}: argument {k} given twice")
            bound[k] = v
        bound = {k: v for k, v in bound.items() if v is not None}
        if family == "Multinomial" and "total_count" in bound:
            # the reference turns every number into a tensor (dist.py:311-318) and torch then refuses it
            raise Exception("Multinomial: inhomogeneous total_count is not supported (leave it at its default of 1, "
                            "as the reference requires)")
        required = [n for n in names if n not in _OPTIONAL and n not in _MATRIX_ARGS
                    and not (family == "Multinomial" and n == "total_count")]
        if family == "MultivariateNormal" and sum(n in bound for n in _MATRIX_ARGS) != 1:
            raise Exception("Exactly one of covariance_matrix or precision_matrix or scale_tril may be specified.")
        for n in required:
            if n not in bound:
                raise Exception(f"Wrong number of arguments provided to {family}")
        if any(n in _OPTIONAL for n in names):
            if sum(n in bound for n in _OPTIONAL) != 1:
                raise Exception(f"{family}: exactly one of probs / logits must be given")
        self.family = family
        self.args = {}
        self.all_args = []
        for k, v in bound.items():
            if isinstance(v, str):
                self.all_args.append(v)
            elif isinstance(v, types.FunctionType):
                self.all_args.extend(function_arguments(v))
            elif isinstance(v, torch.Tensor):
                v = v.detach()
            elif isinstance(v, Param):
                pass                                   # becomes a named parameter when a Problem binds the model (qem.py)
            elif not isinstance(v, numbers.Number):
                raise Exception(f"{family}.{k}: unsupported argument type {type(v)}")
            self.args[k] = v

        # reference dist.py:140-159
        self.qem_dist = any(isinstance(v, QEMParam) for v in self.args.values())
        self.opt_dist = any(isinstance(v, OptParam) for v in self.args.values())
        if self.qem_dist:
            vals = list(self.args.values())
            if not all(isinstance(v, QEMParam) for v in vals) or \
                    any(set(v.ignore_platenames) != set(vals[0].ignore_platenames) for v in vals[1:]):
                raise Exception("If one parameter on a distribution is a QEMParam, then all parameters on that "
                                "distribution should be QEM distributions")

    def __repr__(self):
        return f"{self.family}({', '.join(f'{k}={v!r}' for k, v in self.args.items())})"


def _make(family):
    def ctor(*args, **kwargs):
        return Dist(family, *args, **kwargs)
    ctor.__name__ = family
    return ctor


Normal = _make("Normal")
Bernoulli = _make("Bernoulli")
LogNormal = _make("LogNormal")
Laplace = _make("Laplace")
Exponential = _make("Exponential")
Gamma = _make("Gamma")
Beta = _make("Beta")
Poisson = _make("Poisson")
Cauchy = _make("Cauchy")
HalfNormal = _make("HalfNormal")
Uniform = _make("Uniform")
StudentT = _make("StudentT")
NegativeBinomial = _make("NegativeBinomial")
Binomial = _make("Binomial")
MultivariateNormal = _make("MultivariateNormal")
Dirichlet = _make("Dirichlet")
Gumbel = _make("Gumbel")
Weibull = _make("Weibull")
Pareto = _make("Pareto")
HalfCauchy = _make("HalfCauchy")
Chi2 = _make("Chi2")
Geometric = _make("Geometric")
Kumaraswamy = _make("Kumaraswamy")
FisherSnedecor = _make("FisherSnedecor")
RelaxedBernoulli = _make("RelaxedBernoulli")
OneHotCategorical = _make("OneHotCategorical")
Categorical = _make("Categorical")
ContinuousBernoulli = _make("ContinuousBernoulli")
VonMises = _make("VonMises")
LowRankMultivariateNormal = _make("LowRankMultivariateNormal")
RelaxedOneHotCategorical = _make("RelaxedOneHotCategorical")
Multinomial = _make("Multinomial")


class Data:
    """Marks a variable of Q as observed (reference: src/alan/Data.py)."""
    def __repr__(self):
        return "Data()"


class Timeseries:
    """Timeseries(init, trans) -- reference: src/alan/Timeseries.py:68-87.

    `init` names a variable of the immediately enclosing plate; `trans` is the
    transition distribution, which refers to the previous step as ``prev``
    (Timeseries.py:232-236).
    """
    is_timeseries = True

    def __init__(self, init: str, trans: Dist):
        if not isinstance(init, str):
            raise Exception("the first / `init` argument in a Timeseries should be a string, representing a "
                            "variable name in the above plate")
        if not isinstance(trans, Dist):
            raise Exception("the second / `trans` argument in a Timeseries should be a distribution")
        if trans.qem_dist or trans.opt_dist:
            raise Exception("You can't use QEMParam / OptParam in a timeseries at present")
        self.init = init
        self.trans = trans
        self.all_args = [init, *trans.all_args]


class Group:
    """Variables that share one K axis (reference: src/alan/Group.py)."""
    def __init__(self, **kwargs):
        if len(kwargs) < 2:
            raise Exception("Groups only make sense if they have two or more random variables")
        for v in kwargs.values():
            if not isinstance(v, (Dist, Timeseries)):
                raise Exception("Group members must be distributions or Timeseries")
        self.prog = kwargs


class Plate:
    """Mirror of reference Plate.grouped_prog / flat_prog (src/alan/Plate.py:50-83)."""
    def __init__(self, **kwargs):
        self.grouped_prog = {}
        self.flat_prog = {}
        for k, v in kwargs.items():
            if isinstance(v, Plate):
                self.grouped_prog[k] = v
                self.flat_prog[k] = v
            else:
                if not isinstance(v, (Group, Dist, Timeseries, Data)):
                    raise Exception(f"{k}: a Plate holds distributions, Groups, Data() and sub-Plates")
                group = v.prog if isinstance(v, Group) else {k: v}
                self.grouped_prog[k] = dict(group)
                for gk, gv in group.items():
                    self.flat_prog[gk] = gv
        names = self.all_prog_names()
        dup = sorted({n for n in names if names.count(n) > 1})
        if dup:
            raise Exception(f"Plate has duplicate names {dup}.")

    def all_prog_names(self):
        result = []
        for k, v in self.grouped_prog.items():
            result.append(k)
            if isinstance(v, dict):
                if len(v) >= 2:
                    result.extend(v.keys())
            else:
                result.extend(v.all_prog_names())
        return result

    def groupvarnames(self):
        """Latent groups in program order (each owns one K axis, Plate.py:217-230)."""
        result = []
        for k, v in self.grouped_prog.items():
            if isinstance(v, dict):
                if not datagroup(v):
                    result.append(k)
            else:
                result.extend(v.groupvarnames())
        return result

    def varname2groupvarname(self):
        result = {}
        for k, v in self.grouped_prog.items():
            if isinstance(v, dict):
                if not datagroup(v):
                    for gk in v:
                        result[gk] = k
            else:
                result.update(v.varname2groupvarname())
        return result

    def groupvarname2platenames(self, active=()):
        result = {}
        for k, v in self.grouped_prog.items():
            if isinstance(v, dict):
                if not datagroup(v):
                    result[k] = tuple(active)
            else:
                result.update(v.groupvarname2platenames((*active, k)))
        return result

    def all_platenames(self):
        result = []
        for k, v in self.grouped_prog.items():
            if isinstance(v, Plate):
                result.append(k)
                result.extend(v.all_platenames())
        return result


def datagroup(group: dict) -> bool:
    """reference dist.py:14-19"""
    hasdata = any(isinstance(v, Data) for v in group.values())
    assert not (len(group) >= 2 and hasdata)
    return hasdata


def Kname(groupvarname: str) -> str:
    """Name of the K axis owned by a latent group (reference Plate.py:226: Dim(f"K_{groupname}", K))."""
    return f"K_{groupvarname}"


def check_PQ(P: Plate, Q: Plate, data_names):
    """Structure agreement between P and Q (reference checking.py:56-115, condensed)."""
    if set(P.flat_prog.keys()) != set(Q.flat_prog.keys()):
        raise Exception(f"P and Q must have the same variables/plates at each level; "
                        f"P has {sorted(P.flat_prog)}, Q has {sorted(Q.flat_prog)}")
    for k, q in Q.flat_prog.items():
        p = P.flat_prog[k]
        if isinstance(q, Plate) != isinstance(p, Plate):
            raise Exception(f"{k} is a plate in one of P/Q but not in the other")
        if isinstance(q, Plate):
            check_PQ(p, q, data_names)
        elif isinstance(q, Data):
            if k not in data_names:
                raise Exception(f"{k} is Data() in Q but no data was provided for it")
            if isinstance(p, Data):
                raise Exception(f"{k} is Data() in P; data can only be marked in Q")
        else:
            if k in data_names:
                raise Exception(f"data was provided for {k}, which Q samples")
