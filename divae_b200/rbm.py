"""RBM prior — same API as the reference's models/rbm/rbm.py:14-80, CUDA arithmetic.

Parameters keep the reference's names (`_weights`, `_visible_bias`, `_hidden_bias`),
shapes, dtype (fp32 nn.Parameter, requires_grad) and initialisation (rbm.py:28-34), so
optimisers and state_dict keys are unchanged (SURVEY.md §5 checkpoint).
"""
import logging

import torch
from torch import nn

from . import _ops

logger = logging.getLogger(__name__)


class RBM(nn.Module):
    def __init__(self, n_visible, n_hidden, **kwargs):
        super(RBM, self).__init__(**kwargs)
        self._n_visible = n_visible
        self._n_hidden = n_hidden
        require_grad = True
        # rbm.py:28 — weights ~ N(0, 0.01^2); rbm.py:32 — visible bias 0.5; rbm.py:34 — hidden bias 0
        self._weights = nn.Parameter(torch.randn(n_visible, n_hidden) * 0.01, requires_grad=require_grad)
        self._visible_bias = nn.Parameter(torch.ones(n_visible) * 0.5, requires_grad=require_grad)
        self._hidden_bias = nn.Parameter(torch.zeros(n_hidden), requires_grad=require_grad)
        self._logZ = None          # set by estimate_logZ(); None -> the reference's placeholder
        self._logZ_versions = None
        self._pack = _ops.ParamPack()

    # rbm.py:36-46
    @property
    def visible_bias(self):
        return self._visible_bias

    @property
    def hidden_bias(self):
        return self._hidden_bias

    @property
    def weights(self):
        return self._weights

    # legacy getters used by the reference's dead code (contrastiveDivergence.py:35-37)
    def get_weights(self):
        return self._weights

    def get_visible_bias(self):
        return self._visible_bias

    def get_hidden_bias(self):
        return self._hidden_bias

    def __repr__(self):
        return "RBM: n_vis={0}, n_hid={1}".format(self._n_visible, self._n_hidden)

    def get_logZ_value(self):
        """rbm.py:51-54 returns the literal 33.65; an AIS estimate replaces it once computed."""
        if self._logZ is not None:
            if self._logZ_versions == self._param_versions():
                return self._logZ
            import warnings
            warnings.warn("RBM parameters changed since estimate_logZ(): the stored log Z is stale, "
                          "using the reference's literal 33.65 (call estimate_logZ() again)")
            self._logZ = None
        return 33.65

    def _param_versions(self):
        return tuple((p._version, p.data_ptr()) for p in (self._weights, self._visible_bias, self._hidden_bias))

    def estimate_logZ(self, n_chains=16384, n_betas=10000, seed=0x5EED, betas=None, group=None, return_logw=False,
                      variant="umma"):
        """Annealed-importance-sampling estimate of log Z (new capability; the reference hard-codes
        33.65, rbm.py:51-54).  Chains shard over the ranks of `group` (or the default process group
        when torch.distributed is initialised): every rank runs n_chains / world_size chains keyed by
        their global chain id and the log-weights are combined with an NCCL log-sum-exp.
        `variant`: "umma" anneals with W rounded to bf16 (as the sampler draws); "umma_split" carries the fp32 W as
        bf16 high + low parts (fp32-grade pre-activations, about twice the tensor time).
        Stores the estimate, together with the parameter versions it was computed at, so that get_logZ_value() /
        log_prob() use it until the parameters change (then they fall back to the reference's literal and warn).  Algorithm: include/rbm_b200.h
        (rbm_b200_ais_run), CPU restatement: oracle/rbm_oracle.py:ais_logZ."""
        import ctypes
        import math
        import numpy as np
        import torch.distributed as tdist
        from . import _lib, dist as _dist
        lib = _lib.load()
        _lib.require_cuda(self._weights, "RBM parameters")
        dev = self._weights.device
        rank, world = 0, 1
        if tdist.is_available() and tdist.is_initialized():
            rank, world = tdist.get_rank(group), tdist.get_world_size(group)
        offset, n_local = _dist.shard_chains(n_chains, rank, world)
        if betas is None:
            betas = np.linspace(0.0, 1.0, int(n_betas) + 1)
        betas = np.ascontiguousarray(betas, dtype=np.float64)
        if variant not in ("umma", "umma_split", "auto"):
            raise ValueError("estimate_logZ runs on the UMMA kernels: variant must be 'umma' or 'umma_split'")
        p, keep = self._pack.pack(self)
        logw = torch.empty(n_local, dtype=torch.float64, device=dev)
        nbytes = lib.rbm_b200_ais_workspace(p.n_visible, p.n_hidden, n_local)
        ws = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=dev)
        with torch.no_grad(), torch.cuda.device(dev):
            _lib.check(lib.rbm_b200_ais_run_ex(ctypes.byref(p), n_local, betas.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                               len(betas), int(seed) & 0xFFFFFFFFFFFFFFFF, offset, _ops._ptr(logw),
                                               _lib.VARIANTS[variant], _ops._ptr(ws), nbytes, _ops._stream()), "ais_run")
            lse = _dist.logsumexp_allreduce(logw, group=group)
            a = self._visible_bias.detach().double()
            logZA = p.n_hidden * math.log(2.0) + torch.nn.functional.softplus(a).sum()
            logZ = float((logZA + lse - math.log(n_chains)).item())
        self._logZ = logZ
        self._logZ_versions = self._param_versions()
        if return_logw:
            return logZ, logw
        return logZ

    def energy(self, post_samples):
        """Energy of samples, vis*b_vis + hid*b_hid + vis*w*hid (rbm.py:56-72).

        `post_samples` is a list of [B, *] tensors, concatenated and split in halves
        exactly like the reference.  Differentiable w.r.t. samples and parameters.
        """
        post_samples_concat = torch.cat(post_samples, dim=1)
        n_split = post_samples_concat.size()[1] // 2
        v, h = torch.split(post_samples_concat, split_size_or_sections=int(n_split), dim=1)[:2]
        pack = self._pack.pack(self, need_bf16=False)
        return _ops.energy(self, pack, v, h)

    def cross_entropy(self, post_samples):
        """rbm.py:74-76."""
        return -self.log_prob(post_samples)

    def log_prob(self, samples):
        """log(exp(-E(z))/Z) = -E(z) - log(Z) (rbm.py:78-80)."""
        return -self.energy(samples) - self.get_logZ_value()
