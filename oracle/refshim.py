"""Import the LIVE reference (read-only, /root/reference) with three import-time stubs.

Test infrastructure only; works only in the build container (the reference does not travel to
the GPU box).  Stubs: ``matplotlib.pyplot`` (imported, unused, trainer.py:16),
``soundfile`` + ``LibsndfileError`` (meldataset.py:10-11) and an empty ``pyworld`` so that
``PyWorldBackend`` constructs (f0_backends.py:112-122; it is never called when F0 is supplied).
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("PE_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "model.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load():
    """Returns a namespace with the reference modules: model, trainer, meldataset, optimizers, f0_backends."""
    if not available():
        raise RuntimeError("reference tree not present (expected at %s)" % REF_ROOT)
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot")
    mpl.pyplot = plt

    class LibsndfileError(Exception):
        pass

    _stub("soundfile", LibsndfileError=LibsndfileError, info=None, read=None, SoundFile=None)
    _stub("pyworld")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ns = types.SimpleNamespace()
    for name in ("model", "optimizers", "f0_backends", "meldataset", "trainer"):
        # the reference modules have generic names; load them under a private prefix is not possible
        # (they import each other by bare name), so import as-is from REF_ROOT.
        setattr(ns, name, importlib.import_module(name))
    return ns
