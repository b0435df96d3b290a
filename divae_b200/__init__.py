"""B200-native RBM block-Gibbs sampler: drop-in for DiVAE's models/rbm + models/samplers.

Public surface mirrors the reference (ezeeEric/DiVAE):
    divae_b200.rbm.RBM                          <-> models.rbm.rbm.RBM
    divae_b200.samplers.pcd.PCD                 <-> models.samplers.pcd.PCD
    divae_b200.samplers.gibbsSampler.GibbsSampler <-> models.samplers.gibbsSampler.GibbsSampler
    divae_b200.samplers.baseSampler.BaseSampler <-> models.samplers.baseSampler.BaseSampler
`divae_b200.dropin.install()` registers them under the reference's module paths.
    divae_b200.decoder.DecoderRunner / generate_samples <-> decoder(prior_samples) + shower tail of generate_samples
The arithmetic runs in librbm_b200.so (hand-written sm_100a CUDA behind a C ABI).
"""
from .rbm import RBM  # noqa: F401
from .samplers.baseSampler import BaseSampler  # noqa: F401
from .samplers.gibbsSampler import GibbsSampler  # noqa: F401
from .samplers.pcd import PCD  # noqa: F401
from ._ops import energy_exp  # noqa: F401
from .decoder import DecoderRunner, EncoderRunner, generate_samples  # noqa: F401

__all__ = ["RBM", "BaseSampler", "GibbsSampler", "PCD", "energy_exp", "DecoderRunner", "EncoderRunner", "generate_samples"]
