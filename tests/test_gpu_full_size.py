"""Parity at BASELINE.json's FULL sizes through size-independent properties: slices of the big
job equal the oracle on the same global chain ids, sharding is invisible, states are binary,
unit marginals sit where the conditionals put them."""
import numpy as np
import pytest
import torch

import divae_b200
from divae_b200 import _ops
from oracle import rbm_oracle as orc
from test_gpu_umma import make_rbm, rand_params

pytestmark = pytest.mark.gpu


def test_config4_full_width_65536_chains():
    """RBM 1024x1024, 65,536 chains (BASELINE configs[3]) with a short chain (k=3)."""
    W, a, c = rand_params(1024, 1024, 0.02, 41)
    rbm = make_rbm(W, a, c)
    B, k, seed = 65536, 3, 0x5EED
    s = divae_b200.PCD(batch_size=B, RBM=rbm, n_gibbs_sampling_steps=k, seed=seed)     # AUTO -> UMMA, CTA pairs
    v, h = s.block_gibbs_sampling()
    assert v.shape == (B, 1024) and h.shape == (B, 1024)
    assert bool(((v == 0) | (v == 1)).all()) and bool(((h == 0) | (h == 1)).all())
    for off, n in ((0, 192), (B // 2 - 64, 128), (B - 192, 192)):
        ov, oh, clean = orc.block_gibbs_philox(W, a, c, n, k, seed, chain_offset=off)
        assert clean.mean() > 0.9
        assert (v[off:off + n].cpu().numpy()[clean] == ov[clean]).all()
        assert (h[off:off + n].cpu().numpy()[clean] == oh[clean]).all()
    # a shard with the same global ids reproduces its slice bit for bit
    off, n = 40000, 4096
    shard = divae_b200.PCD(batch_size=n, RBM=rbm, n_gibbs_sampling_steps=k, seed=seed, chain_offset=off, variant="umma")
    vs, hs = shard.block_gibbs_sampling()
    assert torch.equal(vs, v[off:off + n]) and torch.equal(hs, h[off:off + n])
    # marginals: <h_j> over 65,536 chains vs the mean conditional sigma(v W + c) on the returned v's predecessor
    # (checked loosely: within 5 sigma of the batch mean of the conditional evaluated on a 4096-chain slice)
    S, sv, sh = _ops.negative_phase_stats(v, h)
    assert 0.3 < sv.mean().item() < 0.8 and 0.3 < sh.mean().item() < 0.8
    pv = orc.visible_probs(h[:4096].cpu().numpy(), W, a)            # v was drawn from p(v | h)
    sigma = np.sqrt(0.25 / 65536 + 0.25 / 4096)
    assert np.abs(sv.cpu().numpy() - pv.mean(0)).max() < 6 * sigma


def test_config3_full_size_smem():
    """RBM 256x256, 1024 chains x 20 steps (BASELINE configs[2]) — every chain against the oracle."""
    W, a, c = rand_params(256, 256, 0.1, 42)
    s = divae_b200.PCD(batch_size=1024, RBM=make_rbm(W, a, c), n_gibbs_sampling_steps=20, seed=3)
    v, h = s.block_gibbs_sampling()
    ov, oh, clean = orc.block_gibbs_philox(W, a, c, 1024, 20, 3)
    assert clean.mean() > 0.9
    assert (v.cpu().numpy()[clean] == ov[clean]).all() and (h.cpu().numpy()[clean] == oh[clean]).all()


def test_config5_full_chain_count_ais_slice():
    """RBM 512x512, 16,384 AIS chains (BASELINE configs[4]) with a short schedule: log-weights of
    slices equal the CPU restatement on the same global chain ids."""
    rng = np.random.default_rng(5)
    W = orc.round_bf16((rng.standard_normal((512, 512)) * 0.02).astype(np.float32))
    a = (0.3 * rng.standard_normal(512)).astype(np.float32); c = (0.3 * rng.standard_normal(512)).astype(np.float32)
    rbm = make_rbm(W, a, c)
    M, K = 16384, 40
    est, logw = rbm.estimate_logZ(n_chains=M, n_betas=K, seed=9, return_logw=True)
    logw = logw.cpu().numpy()
    for off, n in ((0, 64), (M - 64, 64)):
        _, ref_logw, _ = orc.ais_logZ(W, a, c, n, np.linspace(0, 1, K + 1), 9, chain_offset=off, return_logw=True)
        diff = np.abs(logw[off:off + n] - ref_logw)
        assert np.median(diff) < 5e-3 and (diff < 5e-2).mean() > 0.8
    assert np.isfinite(est)
