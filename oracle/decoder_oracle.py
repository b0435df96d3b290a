"""CPU restatement of the decoder step that follows the sampler in generation.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline leg, never by the product path.

Restates, in numpy fp32 like the reference's ATen ops:
  BasicDecoder.forward    models/networks/basicCoders.py:33-42   (chain, last layer Identity)
  BasicDecoderV3.forward  models/networks/basicCoders.py:68-84   (trunk, two heads, last head layer Identity)
  shower tail             models/autoencoders/gumboltCaloV5.py:91-93 with GumbelMod(is_training=False)
                          utils/dists/gumbelmod.py:18-24
Pinned against the reference's own classes by oracle/gen_golden.py:gen_decoder (tests/golden/decoder_v3.npz,
decoder_basic.npz) and, live, by tests/test_reference_live.py.
"""
import numpy as np

from . import philox
from .rbm_oracle import _step_words, sigmoid

F32 = np.float32
TAG_HITS = 4   # Philox tag of the hit draws (csrc/gibbs_umma.cu: TAG_HITS)


def linear(x, W, b):
    """nn.Linear: x W^T + b (models/networks/networks.py:33-35)."""
    return (np.asarray(x, dtype=F32) @ np.asarray(W, dtype=F32).T + np.asarray(b, dtype=F32)).astype(F32)


def relu(x):
    return np.maximum(x, F32(0))


def chain_forward(layers, x, last_linear=True):
    """layers = [(W, b), ...]; ReLU after every layer except (if last_linear) the last one."""
    n = len(layers)
    for i, (W, b) in enumerate(layers):
        x = linear(x, W, b)
        if not (last_linear and i == n - 1):
            x = relu(x)
    return x


def decoder_basic(layers, x):
    """BasicDecoder.forward (basicCoders.py:33-42) with output_activation_fct = Identity."""
    return chain_forward(layers, x, last_linear=True)


def decoder_v3(trunk, head_hits, head_act, x):
    """BasicDecoderV3.forward (basicCoders.py:68-84): returns (x1 = _layers2 head, x2 = _layers3 head)."""
    t = chain_forward(trunk, x, last_linear=False)
    return chain_forward(head_hits, t, True), chain_forward(head_act, t, True)


def shower_tail(output_hits, output_activations, rho):
    """relu(act) * heaviside(hits + log(rho) - log(1 - rho), 0)   (gumboltCaloV5.py:93, gumbelmod.py:18-24)."""
    rho = np.asarray(rho, dtype=F32)
    logits_gumbel = (np.asarray(output_hits, dtype=F32) + np.log(rho) - np.log(F32(1) - rho)).astype(F32)
    gate = np.where(logits_gumbel > 0, F32(1), F32(0))       # torch.heaviside(x, 0): 0 at x == 0
    return relu(np.asarray(output_activations, dtype=F32)) * gate


def hit_uniforms(seed, n_rows, n_out, chain_offset=0, step=0):
    """u of the device draw contract for the hit gate: unit = output unit, chain = chain_offset + row, TAG_HITS."""
    chains = np.arange(n_rows, dtype=np.uint64) + np.uint64(chain_offset)
    c2, c3 = _step_words(step, TAG_HITS)
    u16 = philox.draw_u16(n_out, chains, c2, c3, seed).astype(np.float64)
    c2e, c3e = _step_words(step, TAG_HITS | philox.TAG_EXTRA_BIT)
    e16 = philox.draw_u16(n_out, chains, c2e, c3e, seed).astype(np.float64)
    return (u16 * 65536.0 + e16) / 4294967296.0


def shower_philox(output_hits, output_activations, seed, chain_offset=0, step=0, tol=1e-4):
    """The tail with the device's own uniforms: gate = [sigmoid(hits) >= u]  (rho = 1 - u in the reference's
    formulation: heaviside(hits + logit(rho)) = [sigmoid(hits) > 1 - rho]).  Returns (samples, clean) where
    clean marks the entries whose draw is at least `tol` away from the threshold."""
    p = sigmoid(np.asarray(output_hits, dtype=np.float64))
    u = hit_uniforms(seed, p.shape[0], p.shape[1], chain_offset, step)
    gate = (p >= u).astype(F32)
    return relu(np.asarray(output_activations, dtype=F32)) * gate, np.abs(p - u) >= tol


# ---------------------------------------------------------------- posterior side
TAG_SMOOTH = 5   # Philox tag of the smoothing draws (csrc/umma_decoder.cu: TAG_SMOOTH)


def smooth_u32(seed, n_rows, n, chain_offset=0, step=0, unit_offset=0):
    """32-bit draws (u16 * 65536 + e16) of the device contract for units unit_offset .. unit_offset + n - 1."""
    chains = np.arange(n_rows, dtype=np.uint64) + np.uint64(chain_offset)
    c2, c3 = _step_words(step, TAG_SMOOTH)
    u16 = philox.draw_u16(unit_offset + n, chains, c2, c3, seed)[:, unit_offset:].astype(np.uint64)
    c2e, c3e = _step_words(step, TAG_SMOOTH | philox.TAG_EXTRA_BIT)
    e16 = philox.draw_u16(unit_offset + n, chains, c2e, c3e, seed)[:, unit_offset:].astype(np.uint64)
    return u16 * np.uint64(65536) + e16


def smooth_rho(u32):
    """rho of the device's training-mode smoothing, in its fp32 arithmetic: 1 - (float(u32) + 0.5) * 2^-32, clamped
    to [2^-24, 1 - 2^-24]."""
    u = (u32.astype(F32) + F32(0.5)) * F32(2.0 ** -32)
    return np.clip(F32(1) - u, F32(2.0 ** -24), F32(1) - F32(2.0 ** -24)).astype(F32)


def gumbel_mod(logits, beta, is_training, rho):
    """GumbelMod.forward (utils/dists/gumbelmod.py:14-24) with an injected rho."""
    logits = np.asarray(logits, dtype=F32); rho = np.asarray(rho, dtype=F32)
    lg = (logits + np.log(rho) - np.log(F32(1) - rho)).astype(F32)
    if is_training:
        return sigmoid(lg * F32(beta))
    return np.where(lg > 0, F32(1), F32(0))


def encoder_forward(levels, x, beta, is_training, rho_of_level):
    """HierarchicalEncoder.forward (models/networks/hierarchicalEncoder.py:139-183) for a tensor input:
    per level  logits = clamp(net(cat([x] + samples)), -88, 88);  samples = GumbelMod(logits, beta, is_training).
    levels = [[(W, b), ...], ...] (ReLU between the layers, last layer linear: hierarchicalEncoder.py:86-96);
    rho_of_level(lvl, shape) supplies the uniforms.  Returns (post_logits, post_samples)."""
    x = np.asarray(x, dtype=F32)
    post_logits, post_samples = [], []
    for lvl, chain in enumerate(levels):
        inp = np.concatenate([x] + post_samples, axis=1)
        logits = np.clip(chain_forward(chain, inp, last_linear=True), F32(-88), F32(88)).astype(F32)
        post_logits.append(logits)
        post_samples.append(gumbel_mod(logits, beta, is_training, rho_of_level(lvl, logits.shape)).astype(F32))
    return post_logits, post_samples


def encoder_forward_philox(levels, x, beta, is_training, seed, chain_offset=0, step=0, tol=1e-4):
    """The same with the device's own draws (level l uses units l * n_latent ...).  Evaluation mode decides
    [sigmoid(l) >= u] against the exact u = u32 / 2^32 like the device; `clean` marks the entries whose draw is at
    least `tol` away from the threshold AT EVERY level that feeds them (a flipped bit changes later levels)."""
    x = np.asarray(x, dtype=F32)
    post_logits, post_samples = [], []
    clean = np.ones(x.shape[0], dtype=bool)
    off = 0
    for chain in levels:
        inp = np.concatenate([x] + post_samples, axis=1)
        logits = np.clip(chain_forward(chain, inp, last_linear=True), F32(-88), F32(88)).astype(F32)
        u32 = smooth_u32(seed, x.shape[0], logits.shape[1], chain_offset, step, off)
        if is_training:
            z = gumbel_mod(logits, beta, True, smooth_rho(u32))
        else:
            p = sigmoid(logits).astype(np.float64)
            u = u32.astype(np.float64) / 4294967296.0
            z = (p >= u).astype(F32)
            clean &= (np.abs(p - u) >= tol).all(axis=1)
        post_logits.append(logits); post_samples.append(z.astype(F32))
        off += logits.shape[1]
    return post_logits, post_samples, clean
