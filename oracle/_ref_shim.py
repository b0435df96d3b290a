"""Import shim for running the real, unmodified reference.

TEST INFRASTRUCTURE.  Makes `from DiVAE import logging` (reference
__init__.py:11-27) and `models.*` / `utils.*` / `engine.*` importable from
the reference tree, and stubs the third-party modules the reference imports but
this image lacks (hydra, torchviz, matplotlib, gif, h5py, coffea — none of them
does arithmetic on the hot path).  The tree is /root/reference in the build container and
the staged copy baseline/_ref/DiVAE (baseline/stage_ref.py; git-ignored, shipped by gpurun)
on the GPU box.
"""
import importlib
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "DiVAE")


def _find_root():
    for cand in (os.environ.get("DIVAE_REFERENCE_ROOT"), "/root/reference", _STAGED):
        if cand and os.path.isfile(os.path.join(cand, "models", "rbm", "rbm.py")):
            return cand
    return "/root/reference"


REF_ROOT = _find_root()


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "models", "rbm", "rbm.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if "DiVAE" not in sys.modules:
        import logging as _logging
        pkg = types.ModuleType("DiVAE")
        pkg.logging = _logging
        pkg.__path__ = [REF_ROOT]
        sys.modules["DiVAE"] = pkg
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    def instantiate(cfg, *args, **kwargs):
        cfg = dict(cfg)
        target = cfg.pop("_target_")
        modname, clsname = target.rsplit(".", 1)
        cls = getattr(importlib.import_module(modname), clsname)
        cfg.update(kwargs)
        return cls(*args, **cfg)

    hydra = _stub("hydra", main=lambda *a, **k: (lambda f: f))
    hydra.utils = _stub("hydra.utils", instantiate=instantiate)
    _stub("torchviz", make_dot=lambda *a, **k: None)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "gif", "h5py", "coffea", "coffea.hist"):
        _stub(name)
    sys.modules["matplotlib.colors"].LogNorm = object
    return REF_ROOT
