"""numpy restatement of the reference's RBM / sampler arithmetic.

TEST INFRASTRUCTURE (see oracle/__init__.py) — the checker, never the product.
Every function cites the reference lines it restates (paths relative to the
reference root, ezeeEric/DiVAE).  Arithmetic is float32 like the reference's
(`torch.float32` parameters and states); states are float32 arrays holding
exactly 0.0 / 1.0.

Pinned by tests/test_oracle_golden.py against tests/golden/*.npz, which were
produced by running the reference's own code (oracle/gen_golden.py).
"""
import numpy as np

from . import philox

F32 = np.float32


def sigmoid(x):
    """torch.sigmoid on float32 (models/samplers/pcd.py:34)."""
    x = np.asarray(x, dtype=F32)
    # numerically stable, evaluated in float32 like ATen's kernel
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = F32(1) / (F32(1) + np.exp(-x[pos], dtype=F32))
    ex = np.exp(x[~pos], dtype=F32)
    out[~pos] = ex / (F32(1) + ex)
    return out


def round_bf16(w):
    """Round float32 -> bfloat16 (round-to-nearest-even) -> float32.

    The CUDA kernels keep W as bf16 (BASELINE.json north_star); parity tests
    hand this rounded W to the oracle so both sides multiply the same numbers.
    """
    w = np.ascontiguousarray(w, dtype=F32)
    bits = w.view(np.uint32).astype(np.uint64)
    rounded = (bits + np.uint64(0x7FFF) + ((bits >> np.uint64(16)) & np.uint64(1))) & np.uint64(0xFFFF0000)
    return rounded.astype(np.uint32).view(F32).reshape(w.shape)


# ---------------------------------------------------------------- RBM (a1-a4)
def rbm_init(n_visible, n_hidden, rng):
    """RBM.__init__ parameter shapes/values (models/rbm/rbm.py:28-34)."""
    W = (rng.standard_normal((n_visible, n_hidden)) * 0.01).astype(F32)
    a = np.full(n_visible, 0.5, dtype=F32)
    c = np.zeros(n_hidden, dtype=F32)
    return W, a, c


def energy(post_samples, W, a, c):
    """RBM.energy (models/rbm/rbm.py:56-72): cat, split in halves, -v.a - h.c - sum((vW)*h)."""
    z = np.concatenate([np.asarray(p, dtype=F32) for p in post_samples], axis=1)
    n_split = z.shape[1] // 2
    v, h = z[:, :n_split], z[:, n_split:2 * n_split]
    e_vis = v @ a
    e_hid = h @ c
    e_mix = np.sum((v @ W) * h, axis=1, dtype=F32)
    return (-e_vis - e_hid - e_mix).astype(F32)


LOGZ_PLACEHOLDER = 33.65  # models/rbm/rbm.py:51-54


def log_prob(samples, W, a, c, logZ=LOGZ_PLACEHOLDER):
    """RBM.log_prob (models/rbm/rbm.py:78-80)."""
    return -energy(samples, W, a, c) - F32(logZ)


# ------------------------------------------------------- PCD half-steps (a7, a8)
def hidden_probs(v, W, c):
    """sigmoid(v W + c) (models/samplers/pcd.py:32-34; gibbsSampler.py:20-23)."""
    return sigmoid(np.asarray(v, dtype=F32) @ W + c)


def visible_probs(h, W, a):
    """sigmoid(h W^T + a) (models/samplers/pcd.py:46-48; gibbsSampler.py:25-28)."""
    return sigmoid(np.asarray(h, dtype=F32) @ W.T + a)


def bernoulli(p, u):
    """(probs >= rand).float() (models/samplers/pcd.py:35,49). u may be float64 (exact)."""
    return (np.asarray(p, dtype=np.float64) >= np.asarray(u, dtype=np.float64)).astype(F32)


def block_gibbs(v0, W, a, c, k, uniforms, tol=1e-6):
    """PCD.block_gibbs_sampling loop body (models/samplers/pcd.py:60-62).

    `uniforms(t, shape)` returns the uniforms of half-step t (t = 2*step for the
    hidden draw, 2*step+1 for the visible draw — the reference's torch.rand call
    order).  Returns (v, h, clean) where clean[b] is False for
    chains that met a near-threshold draw |u-p| < tol anywhere (those may
    legitimately diverge from a different fp32 summation order).
    """
    v = np.asarray(v0, dtype=F32)
    clean = np.ones(v.shape[0], dtype=bool)
    h = None
    for step in range(k):
        ph = hidden_probs(v, W, c)
        u = np.asarray(uniforms(2 * step, ph.shape), dtype=np.float64)
        clean &= ~(np.abs(ph.astype(np.float64) - u) < tol).any(axis=1)
        h = bernoulli(ph, u)
        pv = visible_probs(h, W, a)
        u = np.asarray(uniforms(2 * step + 1, pv.shape), dtype=np.float64)
        clean &= ~(np.abs(pv.astype(np.float64) - u) < tol).any(axis=1)
        v = bernoulli(pv, u)
    return v, h, clean


def init_state_philox(n_chains, n_visible, seed, chain_offset=0, step_offset=0):
    """Device-side replacement of `rand >= rand` (models/samplers/pcd.py:16-17,66-67):
    Bernoulli(1/2) from the top bit of the TAG_INIT field, v0 = 1 - (u16 >> 15)."""
    chains = np.arange(n_chains, dtype=np.uint64) + np.uint64(chain_offset)
    c2, c3 = _step_words(step_offset, philox.TAG_INIT)
    u16 = philox.draw_u16(n_visible, chains, c2, c3, seed)
    return (1 - (u16 >> 15)).astype(F32)


def _step_words(step, tag):
    """64-bit half-step counter -> (c2, c3): low 32 bits, and tag | (high 16 bits << 16)."""
    step = int(step)
    return step & 0xFFFFFFFF, (tag & 0xFFFF) | (((step >> 32) & 0xFFFF) << 16)


def philox_uniforms(seed, n_chains, chain_offset=0, step_offset=0, tag=philox.TAG_GIBBS):
    """uniforms(t, shape) callable implementing the device draw contract (oracle/philox.py)."""
    chains = np.arange(n_chains, dtype=np.uint64) + np.uint64(chain_offset)

    def draw(t, shape):
        c2, c3 = _step_words(step_offset + t, tag)
        u16 = philox.draw_u16(shape[1], chains, c2, c3, seed).astype(np.float64)
        c2e, c3e = _step_words(step_offset + t, tag | philox.TAG_EXTRA_BIT)
        e16 = philox.draw_u16(shape[1], chains, c2e, c3e, seed).astype(np.float64)
        return (u16 * 65536.0 + e16) / 4294967296.0

    return draw


def block_gibbs_philox(W, a, c, n_chains, k, seed, chain_offset=0, step_offset=0, v0=None, tol=1e-6):
    """block_gibbs with the device RNG contract; v0=None draws the initial state on 'device'."""
    if v0 is None:
        v0 = init_state_philox(n_chains, W.shape[0], seed, chain_offset, step_offset)
    return block_gibbs(v0, W, a, c, k, philox_uniforms(seed, n_chains, chain_offset, step_offset), tol)


# -------------------------------------------------- GibbsSampler (a11-a13)
def meanfield_gibbs(x, W, a, c, k):
    """GibbsSampler.gibbs_sampling (models/samplers/gibbsSampler.py:31-37): probabilities, no draw."""
    left = np.asarray(x, dtype=F32)
    right = None
    for _ in range(k):
        right = hidden_probs(left, W, c)
        left = visible_probs(right, W, a)
    return left, right


# ------------------------------------------------- negative phase (a15, F10)
def energy_exp(v, h, W, a, c):
    """GumBolt.energy_exp (models/autoencoders/gumbolt.py:109-134): mean_b(-v^T W h - v.a - h.c)."""
    v = np.asarray(v, dtype=F32); h = np.asarray(h, dtype=F32)
    e = -np.einsum('bi,ij,bj->b', v.astype(np.float64), W.astype(np.float64), h.astype(np.float64))
    e = e - v.astype(np.float64) @ a.astype(np.float64) - h.astype(np.float64) @ c.astype(np.float64)
    return F32(e.mean())


def negative_phase_stats(v, h):
    """Autograd of -energy_exp w.r.t. (W, a, c): <v h^T>, <v>, <h> over the batch."""
    v = np.asarray(v, dtype=np.float64); h = np.asarray(h, dtype=np.float64)
    B = v.shape[0]
    return (v.T @ h / B).astype(F32), (v.mean(axis=0)).astype(F32), (h.mean(axis=0)).astype(F32)


# --------------------------------------------------------------- CD-k (a14)
def cd_k(v0, W, a, c, k, uniforms):
    """Contrastive_Divergence.contrastive_divergence statistics
    (sandbox/rbm_standalone.py:66-73,75-121), without the parameter update.

    `uniforms(i, shape)` -> uniforms of the i-th hidden draw (i = 0 positive phase,
    1..k the Gibbs steps).  The reference draws ONE length-nH vector per draw and
    broadcasts it over the batch (rbm_standalone.py:72,84): pass shape-[1,nH]
    uniforms to restate that, or [B,nH] for independent draws.
    Returns dict(assoc_pos, assoc_neg, dvb, dhb, recon_sse, pv, ph).
    """
    v0 = np.asarray(v0, dtype=F32)
    ph0 = hidden_probs(v0, W, c)                               # :79
    h = bernoulli(ph0, uniforms(0, ph0.shape))                 # :84
    assoc_pos = v0.T @ h                                       # :86
    pv = ph = None
    for i in range(k):                                         # :66-73
        pv = visible_probs(h, W, a)
        ph = hidden_probs(pv, W, c)
        h = bernoulli(ph, uniforms(i + 1, ph.shape))
    assoc_neg = pv.T @ ph                                      # :91
    dvb = np.sum(v0 - pv, axis=0, dtype=F32)                   # :105
    dhb = np.sum(ph0 - ph, axis=0, dtype=F32)                  # :108
    recon = np.sum((v0 - pv) ** 2, dtype=F32)                  # :119-120
    return dict(assoc_pos=assoc_pos.astype(F32), assoc_neg=assoc_neg.astype(F32),
                dvb=dvb, dhb=dhb, recon_sse=recon, pv=pv, ph=ph, ph0=ph0)


def cd_update(state, stats, lr=1e-3, momentum=0.5, weight_cost=1e-4):
    """Momentum / L2 / bias updates (sandbox/rbm_standalone.py:95-115). state: dict W,a,c,dW,da,dc."""
    state['dW'] = state['dW'] * F32(momentum) + (stats['assoc_pos'] - stats['assoc_neg'])
    state['dW'] = state['dW'] - state['W'] * F32(weight_cost)
    state['da'] = state['da'] * F32(momentum) + stats['dvb']
    state['dc'] = state['dc'] * F32(momentum) + stats['dhb']
    state['W'] = state['W'] + state['dW'] * F32(lr)
    state['a'] = state['a'] + state['da'] * F32(lr)
    state['c'] = state['c'] + state['dc'] * F32(lr)
    return state


# ------------------------------------------------- exact enumeration
def _all_states(n):
    idx = np.arange(2 ** n, dtype=np.uint32)
    return ((idx[:, None] >> np.arange(n, dtype=np.uint32)[None, :]) & 1).astype(np.float64)


def exact_logZ(W, a, c):
    """log Z = log sum_h exp(c.h) prod_i (1 + exp(a_i + (W h)_i)); needs min(nV,nH) <= ~22."""
    W = np.asarray(W, dtype=np.float64); a = np.asarray(a, dtype=np.float64); c = np.asarray(c, dtype=np.float64)
    if W.shape[1] > W.shape[0]:
        W, a, c = W.T, c, a
    H = _all_states(W.shape[1])
    act = H @ W.T + a[None, :]
    logterm = H @ c + np.sum(np.logaddexp(0.0, act), axis=1)
    m = logterm.max()
    return float(m + np.log(np.exp(logterm - m).sum()))


def exact_moments(W, a, c):
    """Exact <v>, <h>, <v h^T> of the joint p(v,h) ∝ exp(-E), nV+nH <= ~22."""
    W = np.asarray(W, dtype=np.float64); a = np.asarray(a, dtype=np.float64); c = np.asarray(c, dtype=np.float64)
    nV, nH = W.shape
    V = _all_states(nV); H = _all_states(nH)
    logp = (V @ a)[:, None] + (H @ c)[None, :] + V @ W @ H.T
    logp -= logp.max()
    p = np.exp(logp); p /= p.sum()
    pv = p.sum(axis=1); ph = p.sum(axis=0)
    return V.T @ pv, H.T @ ph, V.T @ p @ H


# --------------------------------------------------------------- AIS (a16)
def softplus(x):
    return np.logaddexp(0.0, x)


def ais_base_bias(a):
    """Base-rate model A: W=0, c=0, visible bias a_A := a."""
    return np.asarray(a, dtype=np.float64)


def ais_logZ(W, a, c, n_chains, betas, seed, chain_offset=0, return_logw=False):
    """Annealed importance sampling (Salakhutdinov & Murray 2008, RBM form).

    The reference has no estimator (models/rbm/rbm.py:51-54 returns the literal 33.65);
    this restates the published algorithm with the device RNG contract:
      v_1 ~ p_A (TAG_AIS_INIT, half_step 0);  for k = 1..K:
        logw += log p*_k(v) - log p*_{k-1}(v);  h ~ sigma(beta_k (vW+c)) (TAG_AIS, half_step 2k);
        v ~ sigma((1-beta_k) a_A + beta_k (hW^T + a)) (TAG_AIS, half_step 2k+1)
      log p*_beta(v) = (1-beta) a_A.v + beta a.v + sum_j softplus(beta (vW+c)_j)
      log Z_A = nH log 2 + sum_i softplus(a_A,i);  log Z = log Z_A + logsumexp(logw) - log M
    With a_A = a the visible-bias terms cancel in the weight increment.
    The last transition (k = K) is skipped: it cannot affect the estimate.
    """
    W32 = np.asarray(W, dtype=F32); a32 = np.asarray(a, dtype=F32); c32 = np.asarray(c, dtype=F32)
    nV, nH = W32.shape
    betas = np.asarray(betas, dtype=np.float64)
    assert betas[0] == 0.0 and betas[-1] == 1.0
    chains = np.arange(n_chains, dtype=np.uint64) + np.uint64(chain_offset)
    aA = a32
    # initial state from the base-rate model
    pA = sigmoid(np.broadcast_to(aA, (n_chains, nV)).copy())
    draw_init = philox_uniforms(seed, n_chains, chain_offset, 0, philox.TAG_AIS_INIT)
    v = bernoulli(pA, draw_init(0, pA.shape))
    draw = philox_uniforms(seed, n_chains, chain_offset, 0, philox.TAG_AIS)
    logw = np.zeros(n_chains, dtype=np.float64)
    K = len(betas) - 1
    for k in range(1, K + 1):
        x = (v @ W32 + c32).astype(F32)                      # one GEMM serves weight and h-draw
        bk, bkm1 = F32(betas[k]), F32(betas[k - 1])
        logw += (softplus((bk * x).astype(np.float64)) - softplus((bkm1 * x).astype(np.float64))).sum(axis=1)
        if k == K:
            break
        ph = sigmoid(bk * x)
        h = bernoulli(ph, draw(2 * k, ph.shape))
        y = (h @ W32.T + a32).astype(F32)
        pv = sigmoid((F32(1) - bk) * aA + bk * y)
        v = bernoulli(pv, draw(2 * k + 1, pv.shape))
    logZA = nH * np.log(2.0) + softplus(aA.astype(np.float64)).sum()
    m = logw.max()
    logZ = float(logZA + m + np.log(np.exp(logw - m).sum()) - np.log(n_chains))
    if return_logw:
        return logZ, logw, float(logZA)
    return logZ
