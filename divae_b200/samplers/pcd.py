"""PCD block-Gibbs sampler — mirrors models/samplers/pcd.py:10-69 on top of librbm_b200.so.

Reference behaviour kept:
  * `PCD(batch_size, RBM, n_gibbs_sampling_steps)`; `get_batch_size()`;
    `hidden_samples(v)`, `visible_samples(h)` draw `(sigmoid(.) >= u).float()`;
  * `block_gibbs_sampling()` returns `(v, h)` float32 {0,1} on the parameters' device,
    h drawn from the previous v and v from that h (pcd.py:60-62,69);
  * the chain state is NOT persistent: every call starts from a fresh Bernoulli(1/2)
    state (pcd.py:64-67) — drawn on the device here instead of CPU RNG + H2D copy;
  * `_RBM` is a registered submodule (duplicate `sampler._RBM.*` state_dict keys),
    `_MCState` is a plain attribute (not a buffer, not checkpointed).
Additive, defaults preserve the reference: `persistent`, `seed`, `variant`,
`chain_offset` (global id of this shard's first chain), `negative_phase()`.
The k-step loop is ONE library call: a persistent shared-memory kernel for small priors,
tcgen05 GEMM half-steps for large ones (BASELINE.json north_star).
"""
import ctypes

import torch

from .. import _lib, _ops
from .baseSampler import BaseSampler


class PCD(BaseSampler):

    def __init__(self, batch_size, RBM, persistent=False, seed=None, variant="auto", chain_offset=None,
                 **kwargs):
        super(PCD, self).__init__(**kwargs)
        self._RBM = RBM
        self.__dict__["_rbm_plain"] = RBM      # same object, reachable without nn.Module.__getattr__ (hot path)
        self._batch_size = batch_size
        # pcd.py:16-17 draws the first state with CPU RNG; here the state is drawn on the
        # device at call time, so nothing is materialised until the first call.
        self._MCState = None
        self._persistent = bool(persistent)
        if seed is None:
            # tie the Philox key to torch's global generator so torch.manual_seed()
            # (scripts/run.py:17) still makes runs reproducible
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        self._seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        if variant not in _lib.VARIANTS:
            raise ValueError("variant must be one of %s" % sorted(_lib.VARIANTS))
        self._variant = _lib.VARIANTS[variant]
        # global id of this sampler's first chain.  None: 0 for a single process; under torch.distributed
        # negative_phase() places the ranks' batches one after the other (rank r starts after the chains of ranks < r)
        self._chain_offset_given = chain_offset is not None
        self._chain_offset = int(chain_offset) if chain_offset is not None else 0
        self._dist_layout = {}   # group id -> (first chain of this rank, total chains over the group)
        self._half_steps = 0  # draw counter: advances by 2k per call
        self._pack = _ops.ParamPack()
        self._workspace = None
        self._ws_query = {}   # (nV, nH, B, variant) -> workspace bytes

    def get_batch_size(self):
        return self._batch_size

    # --------------------------------------------------------------- single half-steps
    def _next_half_step(self):
        t = self._half_steps
        self._half_steps += 1
        return t

    def hidden_samples(self, visible_states):
        """h ~ Bernoulli(sigmoid(v W + c)) (pcd.py:23-35)."""
        pack = self._pack.pack(self._RBM, need_bf16=False)
        with torch.no_grad():
            return _ops.half_step(pack, 0, visible_states, _lib.MODE_SAMPLE, seed=self._seed,
                                  chain_offset=self._chain_offset, half_step_idx=self._next_half_step())

    def visible_samples(self, hidden_states):
        """v ~ Bernoulli(sigmoid(h W^T + a)) (pcd.py:37-49)."""
        pack = self._pack.pack(self._RBM, need_bf16=False)
        with torch.no_grad():
            return _ops.half_step(pack, 1, hidden_states, _lib.MODE_SAMPLE, seed=self._seed,
                                  chain_offset=self._chain_offset, half_step_idx=self._next_half_step())

    # --------------------------------------------------------------------- k-step chain
    def _run_chain(self, v, init_chains, k, step_offset, stats_scale=None):
        lib = _lib.load()
        rbm = self.__dict__.get("_rbm_plain")
        if rbm is None or rbm is not self._modules.get("_RBM"):      # the submodule was replaced after construction
            rbm = self.__dict__["_rbm_plain"] = self._RBM
        B = self._batch_size
        p, keep = self._pack.pack(rbm)     # raises for CPU parameters (no CPU fallback); fp32 only, W is rounded in-library
        dev = keep[0].device
        nV, nH = p.n_visible, p.n_hidden
        if v is None:
            v = torch.empty((B, nV), dtype=torch.float32, device=dev)
        h = torch.empty((B, nH), dtype=torch.float32, device=dev)
        qkey = (nV, nH, B, self._variant, stats_scale is not None)
        nbytes = self._ws_query.get(qkey)
        if nbytes is None:
            query = lib.rbm_b200_block_gibbs_stats_workspace if stats_scale is not None else lib.rbm_b200_block_gibbs_workspace
            nbytes = self._ws_query[qkey] = query(nV, nH, B, self._variant)
        if nbytes and (self._workspace is None or self._workspace.numel() < nbytes
                       or self._workspace.device != dev):
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        ws = self._workspace if nbytes else None
        if stats_scale is not None:
            # chain + negative-phase statistics in one call (fused into the final sweep for small priors)
            S = torch.empty((nV, nH), dtype=torch.float32, device=dev)
            sv = torch.empty(nV, dtype=torch.float32, device=dev)
            sh = torch.empty(nH, dtype=torch.float32, device=dev)
            e = torch.empty((), dtype=torch.float32, device=dev)
            with _ops._on_device(dev):
                _lib.check(lib.rbm_b200_block_gibbs_stats(
                    ctypes.byref(p), _ops._ptr(v), _ops._ptr(h), B, k, 1 if init_chains else 0, self._seed,
                    self._chain_offset, step_offset, self._variant, _ops._ptr(ws), nbytes, float(stats_scale),
                    _ops._ptr(S), _ops._ptr(sv), _ops._ptr(sh), _ops._ptr(e), _ops._stream()), "block_gibbs_stats")
            return v, h, (S, sv, sh), e
        with _ops._on_device(dev):
            _lib.check(lib.rbm_b200_block_gibbs(
                ctypes.byref(p), _ops._ptr(v), _ops._ptr(h), B, k, 1 if init_chains else 0, self._seed,
                self._chain_offset, step_offset, self._variant, _ops._ptr(ws), nbytes, _ops._stream()),
                "block_gibbs")
        return v, h

    def block_gibbs_sampling(self, _stats_scale=None):
        """Block Gibbs sampling (pcd.py:51-69). Returns (visible_states, hidden_states)."""
        k = self.n_gibbs_sampling_steps
        step_offset = self._half_steps
        # (no autograd involved: fresh output tensors filled by the library)
        if self._persistent and self._MCState is not None:
            out = self._run_chain(self._MCState.detach().clone(), False, k, step_offset, _stats_scale)
        else:
            out = self._run_chain(None, True, k, step_offset, _stats_scale)
        self.__dict__["_half_steps"] = step_offset + 2 * k      # plain attribute write, not nn.Module.__setattr__
        if self._persistent:
            self._MCState = out[0]  # true PCD (the line the reference commented out, pcd.py:64)
        return out

    def _block_gibbs_with_uniforms(self, v0, uniforms_h, uniforms_v):
        """Test hook: run k steps against injected uniforms [k,B,nH] / [k,B,nV] (fp32 W, GENERAL kernels)."""
        lib = _lib.load()
        p, keep = self._pack.pack(self._RBM, need_bf16=False)
        v = _ops._f32c(v0, "v0").clone()
        uh = _ops._f32c(uniforms_h, "uniforms_h"); uv = _ops._f32c(uniforms_v, "uniforms_v")
        k, B = uh.shape[0], v.shape[0]
        h = torch.empty((B, p.n_hidden), dtype=torch.float32, device=v.device)
        with torch.cuda.device(v.device):
            _lib.check(lib.rbm_b200_block_gibbs_injected(ctypes.byref(p), _ops._ptr(v), _ops._ptr(h), B, k,
                                                         _ops._ptr(uh), _ops._ptr(uv), _ops._stream()),
                       "block_gibbs_injected")
        return v, h

    def sample(self, n_samples, n_gibbs_sampling_steps=None, first_chain=0):
        """Draw `n_samples` independent chains in ONE launch (fresh Bernoulli(1/2) starts).

        `first_chain` shifts the chain ids: a rank that draws chains [first_chain, first_chain + n_samples) of a
        larger job gets exactly that slice of the unsharded result (generation sharded over GPUs).

        The calorimeter models generate `num_samples` showers by calling block_gibbs_sampling()
        `num_samples // batch_size` times in a Python loop, once per conditioning energy
        (gumboltCaloV5.py:78-96, engineCaloV3.py:115-129); this returns the same
        kind of (v, h) pair for all of them at once.  Chain ids continue after this sampler's own batch,
        so the draws never overlap with block_gibbs_sampling()'s."""
        k = self.n_gibbs_sampling_steps if n_gibbs_sampling_steps is None else int(n_gibbs_sampling_steps)
        saved = (self._batch_size, self._chain_offset, self._workspace)
        try:
            self._batch_size = int(n_samples)
            self._chain_offset = saved[1] + saved[0] + int(first_chain)
            self._workspace = None
            with torch.no_grad():
                v, h = self._run_chain(None, True, k, self._half_steps)
        finally:
            self._batch_size, self._chain_offset, self._workspace = saved
        self._half_steps += 2 * k
        return v, h

    def _group_layout(self, group):
        """(first global chain id of this rank, total chains) for batches laid out rank after rank; one
        all_gather of the batch sizes per (group, batch size), cached."""
        key = (id(group), self._batch_size)
        hit = self._dist_layout.get(key)
        if hit is None:
            dev = self._rbm_plain._weights.device if hasattr(self._rbm_plain, "_weights") else self._RBM.weights.device
            world, rank = torch.distributed.get_world_size(group), torch.distributed.get_rank(group)
            mine = torch.tensor([self._batch_size], dtype=torch.int64, device=dev)
            sizes = [torch.zeros_like(mine) for _ in range(world)]
            torch.distributed.all_gather(sizes, mine, group=group)
            sizes = [int(x.item()) for x in sizes]
            hit = self._dist_layout[key] = (sum(sizes[:rank]), sum(sizes))
        return hit

    # ------------------------------------------------------------- negative phase (a15)
    def negative_phase(self, group=None):
        """Run the chain and return (energy_exp, (S_vh, S_v, S_h)).

        `energy_exp` equals GumBolt.energy_exp(v.detach(), h.detach()) (gumbolt.py:102-134) and
        backpropagates -<v h^T>, -<v>, -<h> into (W, a, c) without the [B,nV,nH] broadcast.
        With a process group the statistics are sum-allreduced over chain shards (NCCL) and scaled by the TOTAL
        chain count; unless `chain_offset` was given, rank r's chains follow those of ranks < r, so the result is
        the statistic of one sum(B_r)-chain job (batch sizes may differ between ranks).
        """
        distributed = group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized())
        total = self._batch_size
        if distributed:
            first, total = self._group_layout(group)
            if not self._chain_offset_given:
                # distinct chains on every rank: without this all ranks would draw chains [0, B) from the same seed
                self._chain_offset = first
        v, h, (S, sv, sh), e = self.block_gibbs_sampling(_stats_scale=1.0 / total)
        if distributed and torch.distributed.get_world_size(group) > 1:
            from .. import dist as _dist
            S, sv, sh = _dist.allreduce_stats(S, sv, sh, group=group)
            e = None    # the local value covers this shard only: recompute from the allreduced statistics
        return _ops.energy_exp_from_stats(self._RBM, S, sv, sh, value=e), (S, sv, sh)
