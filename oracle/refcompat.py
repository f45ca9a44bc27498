"""TEST INFRASTRUCTURE ONLY -- import the reference (alan) in THIS container.

Nothing here is shipped or measured: it exists so that ``tests/golden/make_golden.py``
and the reference-parity tests can run the *unmodified* reference sources where
they lie under /root/reference (which does not exist on the GPU box).

torch 2.11 drifted from the torch the reference was written for; three
sampling/binding-side call sites break.  They are patched at
RUNTIME by replacing attributes of the imported modules -- no reference source
is copied or edited, and nothing on the logPQ arithmetic path is touched:

1. ``PermutationSampler.perm`` (Sampler.py:143-148) builds
   ``TorchDimDist(Uniform, ...)`` whose ``arg_constraints`` is now a property.
   Replacement draws ``rand`` directly (Uniform(0,1).sample() == rand).
2. ``BoundPlate.expand_named`` (BoundPlate.py:17-30) calls ``x.expand()`` with
   zero sizes for 0-d parameters outside any plate.
3. ``Timeseries.sample`` (Timeseries.py:123) calls ``t.stack(list, 0)`` on
   first-class-dim tensors, which functorch.dim now intercepts with a different
   signature.  The module-level name ``t`` inside alan.Timeseries is replaced by a
   thin proxy whose ``stack`` goes through ``functorch.dim.stack``.
"""
import os
import sys
import warnings

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# Where the UNMODIFIED reference package can be imported from: its sources where they lie in the build container,
# else the `pip install --target baseline/_ref /root/reference` copy that __graft_entry__.build() makes there
# (git-ignored: never in history; not gpurun-ignored: it travels to the GPU box).
_CANDIDATES = ["/root/reference/src", os.path.join(_ROOT, "baseline", "_ref")]
REFERENCE_SRC = next((p for p in _CANDIDATES if os.path.isdir(os.path.join(p, "alan"))), _CANDIDATES[0])
REFERENCE_TESTS = "/root/reference/tests"
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "alan"))


def install_reference(verbose=True) -> bool:
    """`pip install --no-index --no-deps --target baseline/_ref` of the reference (from a /tmp copy: the source tree
    is read-only and the build writes egg-info).  Build container only; returns False where /root/reference is absent."""
    import shutil
    import subprocess
    import tempfile
    src = "/root/reference"
    if not os.path.isdir(os.path.join(src, "src", "alan")):
        return False
    target = os.path.join(_ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(target, "alan")):
        return True
    tmp = tempfile.mkdtemp(prefix="alan_ref_")
    try:
        shutil.copytree(src, os.path.join(tmp, "reference"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", target, os.path.join(tmp, "reference")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            print("reference install:", "ok" if r.returncode == 0 else r.stdout[-400:] + r.stderr[-400:])
        return r.returncode == 0
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


_alan = None


def import_reference():
    """Returns the imported, runtime-patched reference package ``alan``."""
    global _alan
    if _alan is not None:
        return _alan
    if not reference_available():
        raise RuntimeError("the reference is neither at /root/reference/src (build container) nor installed under "
                           "baseline/_ref (run __graft_entry__.build() in the build container)")
    for p in (_SHIMS, REFERENCE_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore", message=".*[Nn]amed tensors.*")
    warnings.filterwarnings("ignore", message=".*non-tuple sequence.*")
    import torch as t
    import functorch.dim as fd
    from functorch.dim import Dim
    import alan

    # -- patch 1: PermutationSampler.perm ---------------------------------
    S = sys.modules['alan.Sampler']

    def perm(dims, Kdim):
        _dims = tuple(dims)
        u = t.rand([d.size for d in _dims])[_dims]
        return u.argsort(Kdim).order(Kdim)
    S.PermutationSampler.perm = staticmethod(perm)

    # -- patch 2: BoundPlate.expand_named ---------------------------------
    BP = sys.modules['alan.BoundPlate']
    _orig_expand_named = BP.expand_named

    def expand_named(x, names, all_platesizes):
        names_x = [n for n in x.names if n is not None]
        extra = [all_platesizes[n] for n in names if n not in names_x]
        if len(extra) + x.ndim == 0:
            return x.contiguous().refine_names(*names, *x.names)
        return _orig_expand_named(x, names, all_platesizes)
    BP.expand_named = expand_named

    # -- patch 3: torch.stack on dim-tensors inside alan.Timeseries --------
    TS = sys.modules['alan.Timeseries']

    class _TorchProxy:
        def __getattr__(self, name):
            return getattr(t, name)

        @staticmethod
        def stack(tensors, dim=0):
            if dim == 0 and any(isinstance(x, fd.Tensor) for x in tensors):
                new = Dim("_stack", len(tensors))
                return fd.stack(list(tensors), new).order(new)
            return t.stack(tensors, dim)
    TS.t = _TorchProxy()

    _alan = alan
    return alan
