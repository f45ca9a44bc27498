"""The drop-in boundary exercised END TO END from live reference objects on the GPU:

    prob = alan.Problem(P, Q, data); s = prob.sample(K)               # the unmodified reference, on the CPU
    s.elbo_rws(computation_strategy=B200()).backward()                # its own Sample._elbo, hooked (two lines)
    s.marginals(computation_strategy=B200()) / s._moments_uniform_input(..., computation_strategy=B200())

against the reference's OWN results for the same Sample object on the CPU.  The reference is imported from
/root/reference/src in the build container and from baseline/_ref (its `pip install --target`, made by
__graft_entry__.build(); git-ignored, ships with the snapshot) on the GPU box; skipped where neither exists."""
import pytest
import torch as t

import models
from golden_io import rel_err
from oracle.refcompat import reference_available

pytestmark = [pytest.mark.gpu, pytest.mark.reference,
              pytest.mark.skipif(not reference_available(), reason="needs the reference (/root/reference or baseline/_ref)")]


def _problem(alan, case, seed=0, dtype=t.float32):
    model, inputs_fn, kw, K, moms, joints, N = models.CASES[case]
    t.manual_seed(seed)
    inp = inputs_fn(**kw, seed=seed, dtype=dtype)
    P, Q = model(alan)
    bp = alan.BoundPlate(P, inp['platesizes'], inputs=inp['inputs'])
    bq = alan.BoundPlate(Q, inp['platesizes'], inputs=inp['inputs'],
                         extra_opt_params={k: v.clone() for k, v in inp['params'].items()})
    return alan.Problem(bp, bq, inp['data']), K, moms, joints


@pytest.mark.parametrize("case", list(models.CASES))
def test_reference_sample_through_b200_strategy(case):
    from oracle.refcompat import import_reference
    from alan_b200 import alan_adapter as A
    alan = import_reference()
    from alan.utils import generic_dims, generic_order
    prob, K, moms, joints = _problem(alan, case)
    s = prob.sample(K, reparam=False)
    params = dict(prob.Q._opt_params.to_dict())
    # the reference's own numbers for this very Sample
    ref = s.elbo_rws(computation_strategy=alan.no_checkpoint)
    ref_g = t.autograd.grad(ref, list(params.values()), allow_unused=True) if params else ()
    ref_marg = s.marginals(joints=joints, computation_strategy=alan.no_checkpoint)
    mlist = [((v,), alan.moments.RawMoment(models.MOMENT_FUNCS[f])) for v, f in moms]
    ref_mom = s._moments_uniform_input(mlist)
    remove = A.install_hook(alan)
    try:
        strat = A.B200(device="cuda:0")
        L = s.elbo_rws(computation_strategy=strat)
        assert L.is_cuda and L.ndim == 0
        assert rel_err(L.detach().cpu(), ref.detach()) < 1e-5
        if params:
            g = t.autograd.grad(L, list(params.values()), allow_unused=True)
            for (n, _), a, b in zip(params.items(), g, ref_g):
                if b is None:
                    assert a is None or float(a.abs().max()) == 0.0, n
                    continue
                assert rel_err(a.rename(None).cpu(), b.rename(None)) < 3e-4, n
        assert not s.elbo_nograd(computation_strategy=strat).requires_grad
        marg = s.marginals(joints=joints, computation_strategy=strat)
        for key, w in ref_marg.weights.items():
            dims = generic_dims(w)
            mine = generic_order(marg.weights[key], dims)
            assert rel_err(mine.cpu(), generic_order(w, dims)) < 3e-4, key
        mom = s._moments_uniform_input(mlist, computation_strategy=strat)
        for a, b in zip(mom, ref_mom):
            dims = generic_dims(b)
            assert rel_err(generic_order(a, dims).cpu(), generic_order(b, dims)) < 3e-4
        # a second sample of the same problem reuses the compiled plan
        n = len(strat._cache)
        s2 = prob.sample(K, reparam=False)
        L2 = s2.elbo_rws(computation_strategy=strat)
        assert len(strat._cache) == n
        assert rel_err(L2.detach().cpu(), s2.elbo_rws(computation_strategy=alan.no_checkpoint).detach()) < 1e-5
    finally:
        remove()
    # hook removed: the reference refuses the unknown strategy object again or ignores it, but never reaches the GPU
    assert alan.Sample._elbo.__name__ == "_elbo"
