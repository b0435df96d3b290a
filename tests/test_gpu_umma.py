"""GPU parity of the tcgen05/TMEM (UMMA) variant against the oracle and the reference-made
fixtures.  Every call goes through the C ABI (rbm_b200_block_gibbs, variant = UMMA)."""
import numpy as np
import pytest
import torch

import divae_b200
from oracle import rbm_oracle as orc
from helpers import unpack

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_rbm(W, a, c):
    r = divae_b200.RBM(W.shape[0], W.shape[1])
    with torch.no_grad():
        r._weights.copy_(torch.from_numpy(W)); r._visible_bias.copy_(torch.from_numpy(a))
        r._hidden_bias.copy_(torch.from_numpy(c))
    return r.to(DEV)


def rand_params(nV, nH, scale, seed):
    rng = np.random.default_rng(seed)
    W = orc.round_bf16((rng.standard_normal((nV, nH)) * scale).astype(np.float32))
    a = (0.5 + 0.2 * rng.standard_normal(nV)).astype(np.float32)
    c = (0.2 * rng.standard_normal(nH)).astype(np.float32)
    return W, a, c


def run(W, a, c, B, k, seed, variant="umma", **kw):
    s = divae_b200.PCD(batch_size=B, RBM=make_rbm(W, a, c), n_gibbs_sampling_steps=k, seed=seed, variant=variant, **kw)
    v, h = s.block_gibbs_sampling()
    torch.cuda.synchronize()
    return v.cpu().numpy(), h.cpu().numpy()


@pytest.mark.parametrize("nV,nH", [(64, 64), (128, 256), (256, 128)])
def test_single_step_each_direction(nV, nH):
    """k = 1: h checks the MN-major-B half-step, v the K-major-B half-step."""
    W, a, c = rand_params(nV, nH, 0.3, 1)
    B = 130
    v, h = run(W, a, c, B, 1, seed=5)
    ov, oh, clean = orc.block_gibbs_philox(W, a, c, B, 1, 5)
    assert clean.mean() > 0.95
    assert (h[clean] == oh[clean]).all(), "v->h (MN-major B operand) mismatch"
    assert (v[clean] == ov[clean]).all(), "h->v (K-major B operand) mismatch"


@pytest.mark.parametrize("name", ["pcd_64x64", "pcd_256x256", "pcd_784x128", "pcd_200x200", "pcd_refinit_64x64"])
def test_umma_matches_reference_fixture(golden, name):
    g = golden(name)
    W, a, c = g["W"], g["a"], g["c"]
    B, k, seed = int(g["B"]), int(g["k"]), int(g["seed"])
    nV, nH = W.shape
    v, h = run(W, a, c, B, k, seed)
    _, _, clean = orc.block_gibbs_philox(W, a, c, B, k, seed)
    assert clean.mean() > 0.9
    assert (v[clean] == unpack(g["v"], nV)[clean]).all()
    assert (h[clean] == unpack(g["h"], nH)[clean]).all()


def test_umma_large_prior_multi_tile():
    """1024 x 1024 prior (BASELINE config 4 shape), several M tiles, persistent loop > 1 tile per CTA."""
    W, a, c = rand_params(1024, 1024, 0.03, 7)
    B, k = 640, 3
    v, h = run(W, a, c, B, k, seed=0x5EED)
    ov, oh, clean = orc.block_gibbs_philox(W, a, c, B, k, 0x5EED)
    assert clean.mean() > 0.9
    assert (v[clean] == ov[clean]).all() and (h[clean] == oh[clean]).all()


def test_umma_many_tiles_per_cta_and_statistics():
    """More tiles than SMs: 256 x 512 prior, 40k chains -> 2 x 313 tiles; compare with the general kernel."""
    W, a, c = rand_params(256, 512, 0.1, 9)
    B, k = 40000, 2
    v, h = run(W, a, c, B, k, seed=3)
    v2, h2 = run(W, a, c, B, k, seed=3, variant="general")
    # both follow the same draw contract; they may differ only on near-threshold draws
    assert (v != v2).any(axis=1).mean() < 0.01 and (h != h2).any(axis=1).mean() < 0.01
    ov, oh, clean = orc.block_gibbs_philox(W, a, c, 512, k, 3)
    assert (v[:512][clean] == ov[clean]).all() and (h[:512][clean] == oh[clean]).all()
    ov, oh, clean = orc.block_gibbs_philox(W, a, c, 256, k, 3, chain_offset=B - 256)
    assert (v[-256:][clean] == ov[clean]).all() and (h[-256:][clean] == oh[clean]).all()


@pytest.mark.parametrize("nV,nH,B,k", [(1024, 1024, 4096 + 37, 1), (200, 200, 500, 2), (784, 128, 300, 1),
                                       (64, 64, 128, 1), (130, 700, 20001, 1), (320, 256, 70000, 1)])
def test_umma_statistics_on_the_tensor_cores(nV, nH, B, k):
    """rbm_b200_block_gibbs_stats on the UMMA variant: <vh> is a tcgen05 GEMM V^T H over the bf16 state tiles
    (both operands MN-major, chains split over CTAs), <v>/<h> are column sums of the same tiles.  Binary states
    make the raw counts integers, so they must equal the statistics of the returned (v, h) EXACTLY
    (autograd of GumBolt.energy_exp, gumbolt.py:109-134)."""
    W, a, c = rand_params(nV, nH, 0.05, nV + 3 * nH)
    rbm = make_rbm(W, a, c)
    s = divae_b200.PCD(batch_size=B, RBM=rbm, n_gibbs_sampling_steps=k, seed=29, variant="umma")
    v, h, (S, sv, sh), e = s.block_gibbs_sampling(_stats_scale=1.0)
    vt, ht = v.double(), h.double()
    assert torch.equal(S.double(), vt.t() @ ht)
    assert torch.equal(sv.double(), vt.sum(0)) and torch.equal(sh.double(), ht.sum(0))
    # the samples are the ordinary ones
    v2, h2 = run(W, a, c, B, k, seed=29)
    assert (v.cpu().numpy() == v2).all() and (h.cpu().numpy() == h2).all()
    # scaled statistics + energy expectation through negative_phase()
    s3 = divae_b200.PCD(batch_size=B, RBM=rbm, n_gibbs_sampling_steps=k, seed=29, variant="umma")
    e3, (S3, sv3, sh3) = s3.negative_phase()
    torch.testing.assert_close(S3.double(), vt.t() @ ht / B, rtol=1e-6, atol=1e-7)
    ref = -(((vt @ torch.from_numpy(W).double().to(vt.device)) * ht).sum(1) + vt @ torch.from_numpy(a).double().to(vt.device)
            + ht @ torch.from_numpy(c).double().to(vt.device)).mean()
    assert abs(e3.item() - ref.item()) <= 1e-4 * max(1.0, abs(ref.item()))


def test_umma_persistent_state_and_given_v0():
    W, a, c = rand_params(128, 128, 0.2, 11)
    rbm = make_rbm(W, a, c)
    s = divae_b200.PCD(batch_size=200, RBM=rbm, n_gibbs_sampling_steps=2, seed=21, variant="umma", persistent=True)
    v1, h1 = s.block_gibbs_sampling()
    v2, h2 = s.block_gibbs_sampling()      # continues from v1 with step_offset = 4
    ov1, oh1, c1 = orc.block_gibbs_philox(W, a, c, 200, 2, 21)
    ov2, oh2, c2 = orc.block_gibbs_philox(W, a, c, 200, 2, 21, step_offset=4, v0=ov1)
    ok = c1 & c2
    assert (v1.cpu().numpy()[c1] == ov1[c1]).all()
    assert (v2.cpu().numpy()[ok] == ov2[ok]).all() and (h2.cpu().numpy()[ok] == oh2[ok]).all()


def test_umma_chain_sharding_is_invisible():
    W, a, c = rand_params(192, 320, 0.1, 13)
    vf, hf = run(W, a, c, 512, 3, seed=8)
    for off, n in ((0, 256), (256, 256), (100, 50)):
        vs, hs = run(W, a, c, n, 3, seed=8, chain_offset=off)
        assert (vs == vf[off:off + n]).all() and (hs == hf[off:off + n]).all()


def test_auto_variant_picks_a_tensor_core_path_for_large_batches():
    W, a, c = rand_params(512, 512, 0.05, 15)
    v, h = run(W, a, c, 300, 2, seed=2, variant="auto")
    ov, oh, clean = orc.block_gibbs_philox(W, a, c, 300, 2, 2)
    assert (v[clean] == ov[clean]).all() and (h[clean] == oh[clean]).all()
