"""Statistical parity of the fast variants against exact enumeration (north-star: marginals and <vh>
within 3 sigma) and behaviour under saturating pre-activations."""
import numpy as np
import pytest
import torch

import divae_b200
from divae_b200 import _ops
from oracle import rbm_oracle as orc
from test_gpu_umma import make_rbm, run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", ["smem", "umma", "general"])
def test_stationary_moments_match_exact_enumeration(variant):
    rng = np.random.default_rng(2)
    W = orc.round_bf16((rng.standard_normal((7, 6)) * 0.9).astype(np.float32))
    a = (0.3 * rng.standard_normal(7)).astype(np.float32); c = (0.3 * rng.standard_normal(6)).astype(np.float32)
    ev, eh, evh = orc.exact_moments(W, a, c)
    B = 400000
    s = divae_b200.PCD(batch_size=B, RBM=make_rbm(W, a, c), n_gibbs_sampling_steps=40, seed=1234, variant=variant)
    v, h = s.block_gibbs_sampling()
    S, sv, sh = _ops.negative_phase_stats(v, h)
    worst = 0.0
    for emp, ex in ((sv.cpu().numpy(), ev), (sh.cpu().numpy(), eh), (S.cpu().numpy(), evh)):
        sigma = np.sqrt(np.maximum(ex * (1 - ex), 1e-9) / B)
        worst = max(worst, float((np.abs(emp - ex) / sigma).max()))
    assert worst < 4.0, worst          # 91 statistics: max |z| < 4 (3-sigma rule with a multiple-comparison margin)


@pytest.mark.parametrize("variant", ["smem", "umma", "general"])
def test_saturating_preactivations(variant):
    """|x| up to ~100: sigmoid saturates, e^-x overflows/underflows; decisions must still follow the oracle."""
    rng = np.random.default_rng(3)
    W = orc.round_bf16((rng.standard_normal((64, 64)) * 4.0).astype(np.float32))
    a = (10.0 * rng.standard_normal(64)).astype(np.float32); c = (10.0 * rng.standard_normal(64)).astype(np.float32)
    v, h = run(W, a, c, 200, 6, seed=8, variant=variant)
    ov, oh, clean = orc.block_gibbs_philox(W, a, c, 200, 6, 8)
    assert clean.mean() > 0.9
    assert (v[clean] == ov[clean]).all() and (h[clean] == oh[clean]).all()
    assert np.isfinite(v).all() and np.isfinite(h).all()


def test_persistent_chain_over_many_calls_tracks_one_long_oracle_chain():
    rng = np.random.default_rng(4)
    W = orc.round_bf16((rng.standard_normal((32, 48)) * 0.3).astype(np.float32))
    a = (0.2 * rng.standard_normal(32)).astype(np.float32); c = (0.2 * rng.standard_normal(48)).astype(np.float32)
    s = divae_b200.PCD(batch_size=64, RBM=make_rbm(W, a, c), n_gibbs_sampling_steps=3, seed=5, persistent=True)
    for _ in range(7):
        v, h = s.block_gibbs_sampling()
    # 7 calls x 3 steps == one 21-step chain from the same initial state (draw counters continue)
    ov, oh, clean = orc.block_gibbs_philox(W, a, c, 64, 21, 5)
    assert clean.mean() > 0.8
    assert (v.cpu().numpy()[clean] == ov[clean]).all() and (h.cpu().numpy()[clean] == oh[clean]).all()
