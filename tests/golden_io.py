"""Load golden fixtures (tests/golden/*.pt) into NT dictionaries."""
import os

import torch as t

from alan_b200.named import NT

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TAGS = {"f32": t.float32, "f64": t.float64}


def load(case, tag):
    g = t.load(os.path.join(GOLDEN_DIR, f"{case}_{tag}.pt"), weights_only=False)
    nt = lambda d: {k: NT(v[0], v[1]) for k, v in d.items()}
    g["sample_nt"] = nt(g["sample"])
    g["inputs_params_nt"] = {**nt(g["inputs"]), **nt(g["params"])}
    g["data_nt"] = nt(g["data"])
    return g


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-300)).item()


def tol(tag):
    """north_star: 1e-5 relative in fp32, 1e-10 in fp64"""
    return 1e-5 if tag == "f32" else 1e-10


def elem_err(a, b, floor_frac=1e-3):
    """True element-wise relative error: max_i |a_i - b_i| / max(|b_i|, floor_frac * max|b|) -- the floor keeps
    entries that are (numerically) zero from dividing by nothing; unlike `rel_err` it is not blind to errors in
    entries much smaller than the largest one."""
    a, b = a.double(), b.double()
    den = b.abs().clamp(min=floor_frac * float(b.abs().max().clamp(min=1e-300)))
    return ((a - b).abs() / den).max().item()


def f64_truth(case, g, models, ns, O, joints=()):
    """The case's log-evidence, marginals and moments from the CPU oracle in FLOAT64 on the golden's own inputs
    (upcast): the yardstick for "how wrong is fp32" -- the reference's own fp32 golden has a rounding error against
    it, and the CUDA path is required to be no worse than a small multiple of that (tests/test_gpu_parity.py)."""
    up = lambda d: {k: NT(v.t.double() if v.t.is_floating_point() else v.t, v.axes) for k, v in d.items()}
    P, Q = models.build(case, ns, t.float64)
    sample, ip, data = up(g["sample_nt"]), up(g["inputs_params_nt"]), up(g["data_nt"])
    O.EPS_OF = t.float32              # fp32 semantics (its eps is observable: Appendix A5), float64 arithmetic
    try:
        out = {"elbo": O.elbo(P, Q, sample, ip, data)}
        out["marginals"] = O.marginals(P, Q, sample, ip, data, joints=joints)
        moms = [((v,), models.MOMENT_FUNCS[f]) for v, f in g["moment_specs"]]
        out["moments"] = O.moments(P, Q, sample, ip, data, moms)
    finally:
        O.EPS_OF = None
    return out
