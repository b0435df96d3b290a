"""Mean-field "Gibbs" sampler — mirrors models/samplers/gibbsSampler.py:15-63.

Like the reference it never draws: both conditionals return probabilities and the
sweep is differentiable.  Forward passes run in librbm_b200.so.
"""
import logging

import torch

from .. import _ops
from .baseSampler import BaseSampler

logger = logging.getLogger(__name__)


class GibbsSampler(BaseSampler):
    def __init__(self, RBM, **kwargs):
        super(GibbsSampler, self).__init__(**kwargs)
        self._RBM = RBM  # registered as a submodule like the reference (duplicate state_dict keys)
        self._pack = _ops.ParamPack()

    def hidden_samples(self, probabilities_visible):
        """sigmoid(p_v W + c) (gibbsSampler.py:20-23)."""
        pack = self._pack.pack(self._RBM, need_bf16=False)
        return _ops.meanfield_half_step(self._RBM, pack, 0, probabilities_visible)

    def visible_samples(self, probabilities_hidden):
        """sigmoid(p_h W^T + a) (gibbsSampler.py:25-28)."""
        pack = self._pack.pack(self._RBM, need_bf16=False)
        return _ops.meanfield_half_step(self._RBM, pack, 1, probabilities_hidden)

    def gibbs_sampling(self, input_sample, n_gibbs_sampling_steps):
        """k alternating mean-field sweeps (gibbsSampler.py:31-37)."""
        left = input_sample
        for gibbs_step in range(0, n_gibbs_sampling_steps):
            right = self.hidden_samples(left)
            left = self.visible_samples(right)
        return left, right

    def get_samples(self, approx_post_samples=[], n_latent_nodes=100, n_latent_hierarchy_lvls=4,
                    n_gibbs_sampling_steps=10):
        """gibbsSampler.py:47-63, including the mutable default list and the real-valued
        random start -166*rand+88 (replicated, not silently fixed)."""
        logger.debug("get_samples")
        assert n_latent_hierarchy_lvls % 2 == 0, "Number of hierarchy layers should be even"
        if len(approx_post_samples) < 1:
            for i in range(0, n_latent_hierarchy_lvls // 2):
                approx_post_samples.append(-166 * torch.rand([n_latent_nodes]) + 88)
            approx_post_samples = torch.cat(approx_post_samples).detach()
            approx_post_samples = approx_post_samples.to(self._RBM.visible_bias.device)
        left, right = self.gibbs_sampling(approx_post_samples, n_gibbs_sampling_steps)
        return [left, right]
