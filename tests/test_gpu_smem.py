"""GPU parity of the persistent shared-memory (SMEM) variant: W resident in shared memory,
chain states on chip for all 2k half-steps.  Through the C ABI, against the oracle and the
reference-made fixtures."""
import numpy as np
import pytest
import torch

import divae_b200
from oracle import rbm_oracle as orc
from helpers import unpack
from test_gpu_umma import make_rbm, rand_params, run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["pcd_64x64", "pcd_256x256", "pcd_200x200", "pcd_refinit_64x64"])
def test_smem_matches_reference_fixture(golden, name):
    g = golden(name)
    W, a, c = g["W"], g["a"], g["c"]
    B, k, seed = int(g["B"]), int(g["k"]), int(g["seed"])
    nV, nH = W.shape
    v, h = run(W, a, c, B, k, seed, variant="smem")
    _, _, clean = orc.block_gibbs_philox(W, a, c, B, k, seed)
    assert clean.mean() > 0.9
    assert (v[clean] == unpack(g["v"], nV)[clean]).all()
    assert (h[clean] == unpack(g["h"], nH)[clean]).all()


@pytest.mark.parametrize("nV,nH,B,k", [(16, 16, 5, 3), (32, 32, 128, 40), (100, 37, 50, 4), (256, 64, 33, 3),
                                       (64, 256, 17, 3), (255, 129, 20, 2), (1, 1, 3, 2), (192, 192, 40, 3)])
def test_smem_ragged_shapes(nV, nH, B, k):
    W, a, c = rand_params(nV, nH, 0.3, nV * 1000 + nH)
    v, h = run(W, a, c, B, k, seed=17, variant="smem")
    ov, oh, clean = orc.block_gibbs_philox(W, a, c, B, k, 17)
    assert v.shape == (B, nV) and h.shape == (B, nH)
    assert clean.mean() > 0.8
    assert (v[clean] == ov[clean]).all() and (h[clean] == oh[clean]).all()


def test_smem_throughput_mode_many_chains():
    """Enough chains that one warp owns a chain group and CTAs loop over batches."""
    W, a, c = rand_params(128, 128, 0.2, 3)
    B, k = 148 * 128 * 2 + 77, 3
    v, h = run(W, a, c, B, k, seed=4, variant="smem")
    for off, n in ((0, 300), (B - 300, 300), (148 * 128 - 10, 40)):
        ov, oh, clean = orc.block_gibbs_philox(W, a, c, n, k, 4, chain_offset=off)
        assert (v[off:off + n][clean] == ov[clean]).all() and (h[off:off + n][clean] == oh[clean]).all()


def test_smem_given_v0_and_persistence():
    W, a, c = rand_params(64, 96, 0.3, 5)
    rbm = make_rbm(W, a, c)
    s = divae_b200.PCD(batch_size=70, RBM=rbm, n_gibbs_sampling_steps=3, seed=9, variant="smem", persistent=True)
    v1, h1 = s.block_gibbs_sampling()
    v2, h2 = s.block_gibbs_sampling()
    ov1, oh1, c1 = orc.block_gibbs_philox(W, a, c, 70, 3, 9)
    ov2, oh2, c2 = orc.block_gibbs_philox(W, a, c, 70, 3, 9, step_offset=6, v0=ov1)
    ok = c1 & c2
    assert (v1.cpu().numpy()[c1] == ov1[c1]).all() and (h1.cpu().numpy()[c1] == oh1[c1]).all()
    assert (v2.cpu().numpy()[ok] == ov2[ok]).all() and (h2.cpu().numpy()[ok] == oh2[ok]).all()


def test_smem_chain_sharding_is_invisible():
    W, a, c = rand_params(256, 256, 0.1, 6)
    vf, hf = run(W, a, c, 1024, 5, seed=8, variant="smem")      # BASELINE config 3 shape
    for off, n in ((0, 512), (512, 512), (333, 100)):
        vs, hs = run(W, a, c, n, 5, seed=8, variant="smem", chain_offset=off)
        assert (vs == vf[off:off + n]).all() and (hs == hf[off:off + n]).all()


def test_three_variants_agree_up_to_near_threshold_draws():
    W, a, c = rand_params(256, 256, 0.1, 7)
    outs = [run(W, a, c, 2048, 4, seed=12, variant=var) for var in ("general", "smem", "umma")]
    for v, h in outs[1:]:
        assert (v != outs[0][0]).any(axis=1).mean() < 0.01 and (h != outs[0][1]).any(axis=1).mean() < 0.01


def test_auto_uses_smem_for_small_priors_and_default_pcd_shape():
    """The reference's own sampler: PCD(batch_size=128, n_gibbs_sampling_steps=40) on a 64x64 prior (dvaepp.py:57)."""
    W, a, c = rand_params(64, 64, 0.3, 8)
    v, h = run(W, a, c, 128, 40, seed=2, variant="auto")
    ov, oh, clean = orc.block_gibbs_philox(W, a, c, 128, 40, 2)
    assert clean.mean() > 0.9
    assert (v[clean] == ov[clean]).all() and (h[clean] == oh[clean]).all()


@pytest.mark.parametrize("nV,nH,B,k", [(64, 64, 128, 40), (16, 16, 100, 3), (33, 50, 37, 2), (64, 48, 148 * 128 + 5, 2),
                                       (256, 256, 64, 2), (512, 320, 300, 2)])
def test_negative_phase_statistics_fused_into_the_final_sweep(nV, nH, B, k):
    """rbm_b200_block_gibbs_stats: for priors <= 64x64 the statistics come out of the sampling kernel
    itself; either way they equal the statistics of the returned (v, h) exactly (integer counts)."""
    from divae_b200 import _ops
    W, a, c = rand_params(nV, nH, 0.3, nV + nH)
    rbm = make_rbm(W, a, c)
    s = divae_b200.PCD(batch_size=B, RBM=rbm, n_gibbs_sampling_steps=k, seed=23)
    v, h, (S, sv, sh), _e = s.block_gibbs_sampling(_stats_scale=1.0)
    vn, hn = v.cpu().numpy().astype(np.float64), h.cpu().numpy().astype(np.float64)
    assert (S.cpu().numpy() == vn.T @ hn).all()
    assert (sv.cpu().numpy() == vn.sum(0)).all() and (sh.cpu().numpy() == hn.sum(0)).all()
    # and the samples themselves are the ordinary ones
    s2 = divae_b200.PCD(batch_size=B, RBM=rbm, n_gibbs_sampling_steps=k, seed=23)
    v2, h2 = s2.block_gibbs_sampling()
    assert torch.equal(v, v2) and torch.equal(h, h2)
    # negative_phase(): energy expectation + gradients through the fused statistics
    s3 = divae_b200.PCD(batch_size=B, RBM=rbm, n_gibbs_sampling_steps=k, seed=23)
    e, (S3, sv3, sh3) = s3.negative_phase()
    np.testing.assert_allclose(S3.cpu().numpy(), vn.T @ hn / B, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(e.item(), orc.energy_exp(v.cpu().numpy(), h.cpu().numpy(), W, a, c), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("nV,nH,B,k", [(256, 256, 1024, 20), (200, 200, 500, 2), (130, 70, 301, 1), (65, 256, 4099, 1),
                                       (128, 128, 20000, 2)])
def test_smem_statistics_from_the_final_sweep_tiles(nV, nH, B, k):
    """Priors of 65 - 256 units on the SMEM variant (BASELINE configs[2]: 256 x 256, 1024 chains x 20 steps): the final
    sweep leaves bf16 {0,1} tiles of (v, h) in the workspace and <vh>, <v>, <h> are a tcgen05 GEMM + column sums over
    them - no fp32 SIMT pass over the fp32 outputs.  Raw counts are integers: they equal the statistics of the returned
    states EXACTLY (autograd of GumBolt.energy_exp, gumbolt.py:109-134)."""
    import divae_b200
    from test_gpu_umma import make_rbm, rand_params
    W, a, c = rand_params(nV, nH, 0.1, 7 * nV + nH)
    rbm = make_rbm(W, a, c)
    s = divae_b200.PCD(batch_size=B, RBM=rbm, n_gibbs_sampling_steps=k, seed=31, variant="smem")
    v, h, (S, sv, sh), e = s.block_gibbs_sampling(_stats_scale=1.0)
    vt, ht = v.double(), h.double()
    assert torch.equal(S.double(), vt.t() @ ht)
    assert torch.equal(sv.double(), vt.sum(0)) and torch.equal(sh.double(), ht.sum(0))
    # the samples are those of the plain call
    s2 = divae_b200.PCD(batch_size=B, RBM=rbm, n_gibbs_sampling_steps=k, seed=31, variant="smem")
    v2, h2 = s2.block_gibbs_sampling()
    assert torch.equal(v, v2) and torch.equal(h, h2)
    e3, (S3, sv3, sh3) = divae_b200.PCD(batch_size=B, RBM=rbm, n_gibbs_sampling_steps=k, seed=31, variant="smem").negative_phase()
    torch.testing.assert_close(S3.double(), vt.t() @ ht / B, rtol=1e-6, atol=1e-7)
    Wd = torch.from_numpy(W).double().to(vt.device)
    ref = -(((vt @ Wd) * ht).sum(1) + vt @ torch.from_numpy(a).double().to(vt.device)
            + ht @ torch.from_numpy(c).double().to(vt.device)).mean()
    assert abs(e3.item() - ref.item()) <= 1e-4 * max(1.0, abs(ref.item()))
