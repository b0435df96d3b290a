"""Register this package's classes under the reference's module paths.

After `divae_b200.dropin.install()` the reference's own model code
(`from models.rbm.rbm import RBM`, `from models.samplers.pcd import PCD`,
`_target_: models.samplers.pcd.PCD` in configs/sampler/*.yaml, ...) resolves to the
B200 implementations — the models and engines run unchanged.
Call it before importing any reference `models.autoencoders.*` module.
"""
import importlib
import sys
import types

_MAP = {
    "models.rbm.rbm": "divae_b200.rbm",
    "models.samplers.baseSampler": "divae_b200.samplers.baseSampler",
    "models.samplers.pcd": "divae_b200.samplers.pcd",
    "models.samplers.gibbsSampler": "divae_b200.samplers.gibbsSampler",
}


def install():
    for ref_name, ours in _MAP.items():
        mod = importlib.import_module(ours)
        parts = ref_name.split(".")
        # make sure parent packages exist (the reference's own packages if importable)
        for i in range(1, len(parts)):
            parent = ".".join(parts[:i])
            if parent not in sys.modules:
                try:
                    importlib.import_module(parent)
                except Exception:
                    pkg = types.ModuleType(parent)
                    pkg.__path__ = []
                    sys.modules[parent] = pkg
        sys.modules[ref_name] = mod
        setattr(sys.modules[".".join(parts[:-1])], parts[-1], mod)
    return sorted(_MAP)
