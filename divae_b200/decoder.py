"""Decoder MLP on generated prior samples: the step right after the sampler
in shower / image generation.

Host-side mirror of what the reference does in Python on ATen:

    output_hits, output_activations = self.decoder(prior_samples)             # gumboltCaloV5.py:90
    sample = self._energy_activation_fct(output_activations) \
             * self._hit_smoothing_dist_mod(output_hits, beta, False)          # gumboltCaloV5.py:93

`DecoderRunner` wraps a reference decoder module (BasicDecoder / BasicDecoderV3,
models/networks/basicCoders.py:28-84 - anything exposing `_layers` [, `_layers2`, `_layers3`] of
nn.Linear and a ReLU hidden activation) and runs it through `rbm_b200_decoder_forward`: one tcgen05
GEMM per layer on split-bf16 operands (fp32-grade accuracy), the shower tail fused into the last two.
`generate_samples` mirrors GumBoltCaloV5.generate_samples (gumboltCaloV5.py:78-96) on top of
`PCD.sample` + `DecoderRunner`.  No CPU fallback: CPU modules raise.
"""
import ctypes

import torch

from . import _lib, _ops


def _linear_list(mods, what):
    out = []
    for m in mods:
        if not isinstance(m, torch.nn.Linear) or m.bias is None:
            raise RuntimeError("divae_b200 decoder: %s must be nn.Linear layers with bias (got %r)" % (what, m))
        out.append(m)
    return out


class DecoderRunner:
    """Runs a reference decoder module's forward pass (and the shower tail) on the B200 path."""

    def __init__(self, decoder):
        act = getattr(decoder, "_activation_fct", None)
        if act is not None and not isinstance(act, torch.nn.ReLU):
            raise RuntimeError("divae_b200 decoder: only the ReLU hidden activation of the reference configs "
                               "(configs/model/*.yaml: activation_fct: relu) is implemented, got %r" % (act,))
        out_act = getattr(decoder, "_output_activation_fct", None)
        if out_act is not None and not isinstance(out_act, torch.nn.Identity):
            raise RuntimeError("divae_b200 decoder: output_activation_fct must be Identity (basicCoders.py:29,64), got %r"
                               % (out_act,))
        self._decoder = decoder
        self._trunk = _linear_list(decoder._layers, "_layers")
        l2, l3 = getattr(decoder, "_layers2", None), getattr(decoder, "_layers3", None)
        self._hits = _linear_list(l2, "_layers2") if l2 is not None else []
        self._act = _linear_list(l3, "_layers3") if l3 is not None else []
        if len(self._hits) != len(self._act):
            raise RuntimeError("divae_b200 decoder: the two heads must have the same depth")
        if not self._trunk:
            raise RuntimeError("divae_b200 decoder: BasicDecoderV2-style decoders without a shared trunk are not supported")
        self.in_features = self._trunk[0].in_features
        self.out_features = (self._act[-1] if self._act else self._trunk[-1]).out_features
        self.two_heads = bool(self._act)
        self._desc = self._desc_key = self._keep = None
        self._ws = self._ws_key = None

    def _params(self):
        return [t for m in self._trunk + self._hits + self._act for t in (m.weight, m.bias)]

    def _describe(self):
        """(rbm_b200_decoder, key): ctypes description of the current parameters, rebuilt only when a parameter
        tensor was replaced, moved or modified in place (an optimiser step bumps `_version`)."""
        key = tuple((id(t), t._version, t.data_ptr()) for t in self._params())
        if key == self._desc_key:
            return self._desc, key
        keep = []

        def arr(mods):
            if not mods:
                return ctypes.cast(None, ctypes.POINTER(_lib.Linear))
            a = (_lib.Linear * len(mods))()
            for i, m in enumerate(mods):
                w = _ops._f32c(m.weight.detach(), "decoder weight")
                b = _ops._f32c(m.bias.detach(), "decoder bias")
                keep.extend((w, b))
                a[i] = _lib.Linear(m.in_features, m.out_features, w.data_ptr(), b.data_ptr())
            keep.append(a)
            return a

        d = _lib.Decoder(len(self._trunk), len(self._act), arr(self._trunk), arr(self._hits), arr(self._act))
        self._desc, self._desc_key, self._keep = d, key, keep
        return d, key

    def _run(self, x, want_hits, want_act, want_samples, seed=0, chain_offset=0, step=0):
        lib = _lib.load()
        x = _ops._f32c(x, "decoder input")
        if x.dim() != 2 or x.shape[1] != self.in_features:
            raise RuntimeError("size mismatch: decoder expects [B, %d], got %s" % (self.in_features, tuple(x.shape)))
        B = x.shape[0]
        d, key = self._describe()
        dev = x.device
        new = lambda: torch.empty((B, self.out_features), dtype=torch.float32, device=dev)
        hits = new() if want_hits else None
        act = new() if want_act else None
        samples = new() if want_samples else None
        with torch.no_grad(), _ops._on_device(dev):
            n = lib.rbm_b200_decoder_workspace(ctypes.byref(d), B)
            if n == 0 and B > 0:
                _lib.check(-1, "decoder_workspace")
            # the workspace (and the split parameters staged at its start) is kept across calls
            staged = 0
            if self._ws is not None and self._ws.device == dev and self._ws.numel() >= n:
                staged = 1 if self._ws_key == key else 0
            else:
                self._ws = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
            _lib.check(lib.rbm_b200_decoder_forward(ctypes.byref(d), _ops._ptr(x), B, _ops._ptr(hits), _ops._ptr(act),
                                                    _ops._ptr(samples), seed, chain_offset, step, staged,
                                                    _ops._ptr(self._ws), self._ws.numel(), _ops._stream()),
                       "decoder_forward")
            self._ws_key = key if B > 0 else self._ws_key
        return hits, act, samples

    def forward(self, x):
        """Same return value as `decoder(x)`: a tensor (BasicDecoder) or (output_hits, output_activations)."""
        if self.two_heads:
            hits, act, _ = self._run(x, True, True, False)
            return hits, act
        return self._run(x, False, True, False)[1]

    __call__ = forward

    def shower(self, x, seed, chain_offset=0, step=0):
        """relu(output_activations) * [sigmoid(output_hits) >= u] with Philox draws (gumboltCaloV5.py:93)."""
        if not self.two_heads:
            raise RuntimeError("divae_b200 decoder: shower samples need the two-headed decoder (BasicDecoderV3)")
        return self._run(x, False, False, True, seed, chain_offset, step)[2]


class _Chain:
    """Minimal BasicDecoder-shaped view of one hierarchy level's nn.Sequential for DecoderRunner."""

    def __init__(self, linears):
        self._layers = linears
        self._activation_fct = torch.nn.ReLU()


class EncoderRunner:
    """Inference forward of the reference's HierarchicalEncoder with the Gumbel smoother
    (models/networks/hierarchicalEncoder.py:139-183).

    Per hierarchy level: logits = clamp(net(cat([x] + samples)), -88, 88) - the level's MLP through
    rbm_b200_decoder_forward (tcgen05 GEMMs on split-bf16 operands) - and
    samples = GumbelMod(logits, beta, is_training) through rbm_b200_gumbel_smooth (Philox draws).
    With autograd enabled (training) the same forward runs through autograd Functions whose backward passes are
    tcgen05 GEMMs as well (`_ops.linear`: dx = g W, dW = g^T x on split-bf16 operands; `_ops.gumbel_smooth`:
    beta z (1 - z) and the clamp mask), so the posterior side of a training step stays on the B200 path; under
    torch.no_grad(), and in evaluation mode (hard gates carry no gradient), the staged one-launch-per-layer inference
    path is used."""

    def __init__(self, encoder, beta=None):
        nets = getattr(encoder, "_networks", None)
        if nets is None or len(nets) == 0:
            raise RuntimeError("divae_b200 encoder: expected a HierarchicalEncoder with `_networks`")
        smoother = getattr(encoder, "smoothing_dist_mod", None)
        if smoother is not None and type(smoother).__name__ != "GumbelMod":
            raise RuntimeError("divae_b200 encoder: only the Gumbel smoother (gumbelmod.py) is implemented, got %r" % (smoother,))
        self._levels = []
        self._linears = []
        for net in nets:
            mods = list(net)
            linears = [m for m in mods if isinstance(m, torch.nn.Linear)]
            others = [m for m in mods if not isinstance(m, torch.nn.Linear)]
            if not linears or any(not isinstance(m, (torch.nn.ReLU, torch.nn.Identity)) for m in others):
                raise RuntimeError("divae_b200 encoder: levels must be Linear/ReLU stacks (hierarchicalEncoder.py:86-96)")
            self._levels.append(DecoderRunner(_Chain(linears)))
            self._linears.append(linears)
        if beta is None:
            cfg = getattr(encoder, "_config", None)
            beta = getattr(getattr(cfg, "model", None), "beta_smoothing_fct", None)
        if beta is None:
            raise RuntimeError("divae_b200 encoder: pass beta= (config.model.beta_smoothing_fct)")
        self.beta = float(beta)
        self.n_latent = self._levels[0].out_features

    def forward(self, x, is_training=False, seed=0, chain_offset=0, step=0):
        """Returns (beta, post_logits, post_samples) like HierarchicalEncoder.forward (hierarchicalEncoder.py:183)."""
        lib = _lib.load()
        scale = None
        if isinstance(x, tuple):          # (data, per-sample factor): samples are multiplied by x[1] (hierarchicalEncoder.py:178-179)
            x, scale = x
        x = _ops._f32c(x, "encoder input")
        B = x.shape[0]
        post_logits, post_samples = [], []
        if is_training and torch.is_grad_enabled() and (
                x.requires_grad or any(p.requires_grad for ls in self._linears for m in ls for p in m.parameters())):
            # training: differentiable forward, every product (forward and backward) a tcgen05 GEMM
            for lvl, linears in enumerate(self._linears):
                cur = torch.cat([x] + post_samples, dim=1) if post_samples else x
                for i, m in enumerate(linears):
                    cur = _ops.linear(cur, m.weight, m.bias, relu=(i + 1 < len(linears)))
                logits, samples = _ops.gumbel_smooth(cur, self.beta, is_training, seed, chain_offset, step, lvl * self.n_latent)
                if scale is not None:
                    samples = samples * _ops._f32c(scale, "scale").reshape(B, 1)
                post_logits.append(logits); post_samples.append(samples)
            return torch.tensor(self.beta, dtype=torch.float, device=x.device), post_logits, post_samples
        with torch.no_grad(), _ops._on_device(x.device):
            for lvl, runner in enumerate(self._levels):
                inp = torch.cat([x] + post_samples, dim=1) if post_samples else x
                logits = runner(inp)
                samples = torch.empty_like(logits)
                _lib.check(lib.rbm_b200_gumbel_smooth(_ops._ptr(logits), _ops._ptr(samples), B, logits.shape[1], self.beta,
                                                      1 if is_training else 0, 1, seed, chain_offset, step,
                                                      lvl * self.n_latent, _ops._stream()), "gumbel_smooth")
                if scale is not None:
                    samples = samples * _ops._f32c(scale, "scale").reshape(B, 1)
                post_logits.append(logits); post_samples.append(samples)
        beta = torch.tensor(self.beta, dtype=torch.float, device=x.device)
        return beta, post_logits, post_samples

    __call__ = forward


def generate_samples(model, num_samples=64, true_energy=None, seed=0, group=None, gather=False):
    """Drop-in for GumBoltCaloV5.generate_samples (models/autoencoders/gumboltCaloV5.py:78-96).

    The reference loops `num_samples // batch_size` times over block_gibbs_sampling() + decoder; here the chains
    of all iterations are drawn in ONE sampler launch (`PCD.sample`) and decoded in one pass.  Returns
    (true_energies [n,1], samples [n, n_out]) with n = max(num_samples // batch_size, 1) * batch_size.

    With torch.distributed initialised (or `group` given) the n showers are sharded over the ranks as contiguous
    blocks of global shower ids (dist.shard_chains): every rank returns ITS block (or, with gather=True, all of
    them); because the sampler's and the hit gate's draws are keyed by the global id, the union of the blocks is
    what a single GPU would have generated.  Random conditioning energies (true_energy=None) are drawn per rank from
    torch's generator and are therefore not part of that guarantee."""
    sampler = model.sampler
    bs = sampler.get_batch_size()
    n = max(num_samples // bs, 1) * bs
    off, n_local = 0, n
    distributed = group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized())
    if distributed:
        from . import dist as _dist
        rank, world = torch.distributed.get_rank(group), torch.distributed.get_world_size(group)
        off, n_local = _dist.shard_chains(n, rank, world)
    v, h = sampler.sample(n_local, first_chain=off)
    if true_energy is None:
        true_e = torch.rand((n_local, 1), device=v.device) * 100.
    else:
        true_e = torch.ones((n_local, 1), device=v.device) * true_energy
    # one runner per model: it keeps the split parameters staged on the device between calls
    runner = getattr(model, "_b200_decoder_runner", None)
    if runner is None or runner._decoder is not model.decoder:
        runner = DecoderRunner(model.decoder)
        object.__setattr__(model, "_b200_decoder_runner", runner)
    x = torch.cat([v, h, true_e], dim=1)
    samples = runner.shower(x, seed=seed, chain_offset=off)
    if distributed and gather:
        world = torch.distributed.get_world_size(group)
        sizes = [_dist.shard_chains(n, r, world)[1] for r in range(world)]
        outs_e = [torch.empty((m, 1), device=v.device) for m in sizes]
        outs_s = [torch.empty((m, samples.shape[1]), device=v.device) for m in sizes]
        torch.distributed.all_gather(outs_e, true_e.contiguous(), group=group)
        torch.distributed.all_gather(outs_s, samples.contiguous(), group=group)
        return torch.cat(outs_e, 0), torch.cat(outs_s, 0)
    return true_e, samples
