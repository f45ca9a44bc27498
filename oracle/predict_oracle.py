"""TEST INFRASTRUCTURE ONLY -- CPU oracle for prediction.

Plain-torch restatement, with EXPLICIT base noise, of
  extend          ImportanceSample.extend -> Plate.sample_extended -> Dist.sample_extended
                  src/alan/ImportanceSample.py:43-98, Plate.py:145-182, dist.py:234-269
  predictive_ll   ExtendedImportanceSample.predictive_ll -> Plate.predictive_ll -> Dist.predictive_ll
                  src/alan/ImportanceSample.py:118-177, Plate.py:184-215, dist.py:271-294

Noise contract (the keys alan_b200.predict.Extender.noise_shapes() lists): noise[varname] = standard normals / uniforms
of shape [extended plates of the variable..., N, *event]; a draw is the closed-form transform of oracle/sample_oracle.py.

Parity pin: tests/golden/predict_*.pt hold the outputs of the UNMODIFIED reference (`ImportanceSample.extend`,
`ExtendedImportanceSample.predictive_ll`) with `TorchDimDist.sample` replaced by the same explicit-noise transform
(tests/golden/make_golden_predict.py); tests/test_predict_cpu.py checks this file against them.
"""
from __future__ import annotations

import torch as t

from alan_b200.model import Plate, Timeseries
from alan_b200.named import NT
from .logpq_oracle import ONT, dist_log_prob, logmeanexp
from .sample_oracle import draw


def _walk(P: Plate, active=()):
    for name, child in P.flat_prog.items():
        if isinstance(child, Plate):
            yield from _walk(child, (*active, name))
        else:
            yield name, child, tuple(active)


def extend(P: Plate, samples: dict, data: dict, ext_inputs: dict, noise: dict, N: int, dtype=t.float32) -> dict:
    """{varname: NT[plates..., N, *event]} for EVERY variable of P (latents and data), extended."""
    scope = {k: ONT(v.t.to(dtype), v.axes) for k, v in (ext_inputs or {}).items()}
    out = {}
    for var, d, active in _walk(P):
        if isinstance(d, Timeseries):
            raise Exception("oracle: extend through a Timeseries is not restated")
        axes = tuple(active) + ('N',)
        eps = ONT(noise[var].to(dtype), axes)
        x = draw(d, scope, eps, dtype)                                   # [ext plates..., N, *event]
        orig = samples.get(var, data.get(var))
        if orig is not None:
            o = ONT(orig.t.to(dtype), orig.axes)
            o = o.order(tuple(a for a in axes if a in o.axes))
            xt = x.t.clone()
            idx = []
            src = o.t
            for i, a in enumerate(axes):
                if a in o.axes:
                    idx.append(slice(0, o.sizes()[a]))
                else:
                    idx.append(slice(None))
                    src = src.unsqueeze(i)
            xt[tuple(idx)] = src.expand_as(xt[tuple(idx)])
            x = ONT(xt, axes)
        scope[var] = x
        out[var] = NT(x.t, x.axes)
    return out


def predictive_ll(P: Plate, ext_samples: dict, ext_data: dict, orig_sizes: dict, ext_inputs: dict, N: int, dtype=t.float32) -> dict:
    """{data variable: 0-d tensor} = logmeanexp_N( sum_all ll - sum_train ll )."""
    scope = {k: ONT(v.t.to(dtype), v.axes) for d in (ext_inputs or {}, ext_samples) for k, v in d.items() if k not in ext_data}
    out = {}
    for var, d, active in _walk(P):
        if var not in ext_data:
            continue
        val = ONT(ext_data[var].t.to(dtype), ext_data[var].axes)
        ll = dist_log_prob(d, val, scope, dtype)                         # named axes: plates..., N (some order)
        plates = tuple(a for a in ll.axes if a != 'N')
        ll = ll.order(plates + ('N',))
        train = ll.t[tuple(slice(0, orig_sizes[a]) for a in plates)]
        diff = ONT(ll.t.sum(tuple(range(len(plates)))) - train.sum(tuple(range(len(plates)))), ('N',))
        out[var] = logmeanexp(diff, ('N',)).t
    return out
