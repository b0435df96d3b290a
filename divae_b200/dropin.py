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


def accelerate(model, seed=0):
    """Route the rows NEXT to the sampler of an (unchanged) reference model instance through the B200 path as well
    (SURVEY.md section 8f ranks 1 and 4), forward and backward:
      * `model.energy_exp(v, h)` (models/autoencoders/gumbolt.py:109-134: W broadcast to [B, nV, nH]) -> the fused energy
        kernel with tcgen05 backward GEMMs (divae_b200.energy_exp); same value, same gradients;
      * `model.encoder(x, is_training)` (models/networks/hierarchicalEncoder.py:139-183 + GumbelMod) -> EncoderRunner: every
        Linear forward and backward a split-bf16 tcgen05 GEMM, Gumbel noise from the Philox contract (one draw step per call).
    Parameters stay the model's own nn.Parameters (optimisers, checkpoints and state_dict keys unchanged).  Returns the model.
    Pieces that do not have the expected shape (another smoother, non-MLP levels) are left as they are."""
    import itertools
    import torch
    from . import _ops
    from .decoder import EncoderRunner

    prior = getattr(model, "prior", None)
    if prior is not None and hasattr(prior, "_pack"):
        object.__setattr__(model, "energy_exp", lambda rbm_vis, rbm_hid: _ops.energy_exp(prior, rbm_vis, rbm_hid))
    enc = getattr(model, "encoder", None)
    if enc is not None:
        try:
            runner = EncoderRunner(enc)
        except RuntimeError:
            runner = None
        if runner is not None:
            steps = itertools.count()

            def forward(x, is_training=True):
                flat = x if isinstance(x, tuple) else x.reshape(x.shape[0], -1)
                return runner(flat, is_training=is_training, seed=seed, step=next(steps))

            object.__setattr__(enc, "forward", forward)
            object.__setattr__(model, "_b200_encoder_runner", runner)
    return model
