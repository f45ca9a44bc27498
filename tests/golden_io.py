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
