"""TEST INFRASTRUCTURE ONLY -- minimal ``opt_einsum`` stand-in for running the reference here.

The reference calls ``opt_einsum.contract_path(*args)[0]`` once
(/root/reference/src/alan/reduce_Ks.py:265) to choose the ORDER of pairwise
contractions; no arithmetic goes through it.  opt_einsum (unpinned in
/root/reference/setup.py:12) is not installed and there is no network, so the
golden-vector generator puts this directory on sys.path.  It forwards to the
same deterministic greedy rule the product planner uses, so reference, oracle
and CUDA engine all walk the same contraction steps.
"""
from alan_b200.path import greedy_path


def contract_path(*args, **kwargs):
    *pairs, out_idxs = args
    tensors, idxs = pairs[0::2], pairs[1::2]
    sizes = {}
    for ten, ix in zip(tensors, idxs):
        for extent, i in zip(ten.shape, ix):
            sizes[i] = int(extent)
    all_idx = set().union(*[set(ix) for ix in idxs]) if idxs else set()
    sum_axes = all_idx - set(out_idxs)
    return greedy_path(idxs, sum_axes, sizes), None
