"""QEM on CPU: parameter binding (OptParam / QEMParam -> named parameters, initial mean
parameters) and the oracle's update (oracle/qem_oracle.py) against the states the UNMODIFIED reference goes through
(tests/golden/make_golden_qem.py).  The GPU runs the same updates through the engine and the C ABI
(tests/test_gpu_qem.py)."""
import os

import pytest
import torch as t

import models
from alan_b200 import model as M
from alan_b200.named import NT
from alan_b200.qem import bind, SUFFICIENT, CONV_ARGS
from golden_io import GOLDEN_DIR, TAGS, elem_err

CASES = list(models.QEM_CASES)


def load(case, tag):
    return t.load(os.path.join(GOLDEN_DIR, f"{case}_{tag}.pt"), weights_only=False)


def nts(d):
    return {k: NT(v[0].clone(), v[1]) for k, v in d.items()}


def build(case, dtype):
    old = t.get_default_dtype()
    t.set_default_dtype(dtype)
    try:
        return models.QEM_CASES[case][0](M)
    finally:
        t.set_default_dtype(old)


def bound(case, tag):
    g = load(case, tag)
    P, Q = build(case, TAGS[tag])
    taken = set(g['params'])
    Pb, optP, qpP, qmP, qvP = bind(P, g['platesizes'], taken)
    Qb, optQ, qpQ, qmQ, qvQ = bind(Q, g['platesizes'], taken | set(optP) | set(qpP))
    cast = lambda d: {k: NT(v.t.to(TAGS[tag]), v.axes) for k, v in d.items()}
    return g, Pb, Qb, {**cast(optP), **cast(optQ)}, {'P': (qvP, cast(qpP), cast(qmP)), 'Q': (qvQ, cast(qpQ), cast(qmQ))}


def check_state(side_state, ref, tag, what):
    tol = 2e-5 if tag == 'f32' else 1e-9
    _, params, means = side_state
    assert set(params) == set(ref['params']) and set(means) == set(ref['means']), what
    for k, (x, axes) in ref['params'].items():
        assert elem_err(params[k].order(axes).t.cpu(), x) < tol, (what, 'param', k)
    for k, (x, axes) in ref['means'].items():
        assert elem_err(means[k].order(axes).t.cpu(), x) < tol, (what, 'mean', k)


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_binding_matches_boundplate(case, tag):
    """Names, plate expansion and initial mean parameters (conv2mean) equal the reference's BoundPlate."""
    g, Pb, Qb, opt, sides = bound(case, tag)
    assert set(opt) == set(g['opt_params']) - set(g['params'])
    for k in opt:
        x, axes = g['opt_params'][k]
        assert opt[k].axes == axes and t.equal(opt[k].t, x.to(TAGS[tag])), k
    for side in 'PQ':
        check_state(sides[side], g['states'][0][side], tag, f'initial {side}')
    for d in list(Qb.flat_prog.values()) + list(Pb.flat_prog.values()):
        if isinstance(d, M.Dist):
            assert not any(isinstance(v, M.Param) for v in d.args.values())


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("case", CASES)
def test_oracle_update_matches_reference(case, tag):
    from oracle import qem_oracle as QO
    g, Pb, Qb, opt, sides = bound(case, tag)
    sample, data = nts(g['sample']), nts(g['data'])
    spec = lambda qvs: [(q.varname, q.family, q.arg2param, q.meannames) for q in qvs]
    for step, lr in enumerate(g['lrs'], start=1):
        for side in 'PQ':                                   # P first; Q's moments see P's new parameters
            qvs, params, means = sides[side]
            ip = {**nts(g['params']), **opt, **sides['P'][1], **sides['Q'][1]}
            QO.update_side(spec(qvs), params, means, lr, Pb, Qb, sample, ip, data)
        for side in 'PQ':
            check_state(sides[side], g['states'][step][side], tag, f'after update {step} ({side})')


def test_declaration_errors_follow_the_reference():
    with pytest.raises(Exception, match="all parameters on that distribution should be QEM"):
        M.Normal(M.QEMParam(0.), 1.)
    with pytest.raises(Exception, match="timeseries"):
        M.Timeseries('init', M.Normal(M.OptParam(0.), 1.))
    Q = M.Plate(a=M.Cauchy(M.QEMParam(0.), M.QEMParam(1.)))
    with pytest.raises(Exception, match="no mean <-> conventional"):
        bind(Q, {})
    Q = M.Plate(a=M.Normal(M.OptParam(0., name='w'), 1.), b=M.Normal(M.OptParam(0., name='w'), 1.))
    with pytest.raises(Exception, match="already a parameter with this name"):
        bind(Q, {})


def test_family_tables_are_consistent():
    from alan_b200.runtime import QEM_FAMILY
    assert set(SUFFICIENT) == set(CONV_ARGS) == set(QEM_FAMILY)
