"""GPU parity of the fp32 GENERAL kernels against the oracle and the reference-made fixtures.
Everything goes through the C ABI (ctypes) via the host classes."""
import numpy as np
import pytest
import torch

import divae_b200
from divae_b200 import _lib, _ops
from oracle import philox, rbm_oracle as orc
from helpers import unpack

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_rbm(W, a, c):
    r = divae_b200.RBM(W.shape[0], W.shape[1])
    with torch.no_grad():
        r._weights.copy_(torch.from_numpy(W)); r._visible_bias.copy_(torch.from_numpy(a))
        r._hidden_bias.copy_(torch.from_numpy(c))
    return r.to(DEV)


def test_library_sees_a_blackwell_device():
    import ctypes
    lib = _lib.load()
    sm, smem, cc = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    _lib.check(lib.rbm_b200_device_info(ctypes.byref(sm), ctypes.byref(smem), ctypes.byref(cc)))
    assert cc.value >= 100 and sm.value >= 100 and smem.value >= 200 * 1024


@pytest.mark.parametrize("name", ["pcd_64x64", "pcd_256x256", "pcd_784x128", "pcd_200x200", "pcd_refinit_64x64"])
def test_block_gibbs_general_matches_reference_fixture(golden, name):
    """Free-running Philox chain == what the reference returned when fed the same uniforms."""
    g = golden(name)
    W, a, c = g["W"], g["a"], g["c"]
    B, k, seed = int(g["B"]), int(g["k"]), int(g["seed"])
    nV, nH = W.shape
    s = divae_b200.PCD(batch_size=B, RBM=make_rbm(W, a, c), n_gibbs_sampling_steps=k, seed=seed, variant="general")
    v, h = s.block_gibbs_sampling()
    assert v.dtype == torch.float32 and v.shape == (B, nV) and h.shape == (B, nH) and v.is_cuda
    _, _, clean = orc.block_gibbs_philox(W, a, c, B, k, seed)
    v, h = v.cpu().numpy(), h.cpu().numpy()
    assert set(np.unique(v)) <= {0.0, 1.0}
    assert clean.mean() > 0.9
    assert (v[clean] == unpack(g["v"], nV)[clean]).all()
    assert (h[clean] == unpack(g["h"], nH)[clean]).all()


def test_init_chains_matches_contract(golden):
    import ctypes
    lib = _lib.load()
    for B, nV, off, step in ((100, 64, 0, 0), (33, 200, 1000, 46), (5, 784, 2 ** 33, 2 ** 34 + 5)):
        v = torch.empty(B, nV, device=DEV)
        _lib.check(lib.rbm_b200_init_chains(_ops._ptr(v), B, nV, 0x5EED, off, step, _ops._stream()))
        ref = orc.init_state_philox(B, nV, 0x5EED, chain_offset=off, step_offset=step)
        assert (v.cpu().numpy() == ref).all()


def test_half_steps_injected_uniforms_bit_exact_outside_band(golden):
    """PCD.hidden_samples / visible_samples against injected fp32 uniforms (north-star rule:
    bit-exact except where |u - p| < 1e-6)."""
    g = golden("pcd_200x200")
    W, a, c = g["W"], g["a"], g["c"]
    rng = np.random.default_rng(0)
    B = 77
    v = (rng.random((B, 200)) < 0.5).astype(np.float32)
    U = rng.random((B, 200)).astype(np.float32)
    pack = _ops.ParamPack().pack(make_rbm(W, a, c), need_bf16=False)
    h = _ops.half_step(pack, 0, torch.from_numpy(v).to(DEV), _lib.MODE_INJECTED, uniforms=torch.from_numpy(U).to(DEV))
    p = orc.hidden_probs(v, W, c)
    ok = np.abs(p - U) >= 1e-6
    assert ok.mean() > 0.999
    assert (h.cpu().numpy()[ok] == orc.bernoulli(p, U)[ok]).all()
    v2 = _ops.half_step(pack, 1, h, _lib.MODE_INJECTED, uniforms=torch.from_numpy(U).to(DEV))
    p2 = orc.visible_probs(h.cpu().numpy(), W, a)
    ok = np.abs(p2 - U) >= 1e-6
    assert (v2.cpu().numpy()[ok] == orc.bernoulli(p2, U)[ok]).all()


def test_block_gibbs_injected_chain():
    rng = np.random.default_rng(1)
    nV, nH, B, k = 96, 40, 50, 6
    W = (rng.standard_normal((nV, nH)) * 0.2).astype(np.float32)
    a = (0.2 * rng.standard_normal(nV)).astype(np.float32); c = (0.2 * rng.standard_normal(nH)).astype(np.float32)
    v0 = (rng.random((B, nV)) < 0.5).astype(np.float32)
    Uh = rng.random((k, B, nH)).astype(np.float32); Uv = rng.random((k, B, nV)).astype(np.float32)
    s = divae_b200.PCD(batch_size=B, RBM=make_rbm(W, a, c), n_gibbs_sampling_steps=k, seed=1, variant="general")
    v, h = s._block_gibbs_with_uniforms(torch.from_numpy(v0).to(DEV), torch.from_numpy(Uh).to(DEV),
                                        torch.from_numpy(Uv).to(DEV))
    ov, oh, clean = orc.block_gibbs(v0, W, a, c, k, lambda t, shape: (Uh if t % 2 == 0 else Uv)[t // 2])
    assert clean.mean() > 0.95
    assert (v.cpu().numpy()[clean] == ov[clean]).all() and (h.cpu().numpy()[clean] == oh[clean]).all()


def test_probabilities_and_meanfield_within_1e3_relative(golden):
    g = golden("meanfield")
    W, a, c = g["W"], g["a"], g["c"]
    rbm = make_rbm(W, a, c)
    gs = divae_b200.GibbsSampler(RBM=rbm, n_gibbs_sampling_steps=40)
    x = torch.from_numpy(g["x"]).to(DEV)
    left, right = gs.gibbs_sampling(x, int(g["k"]))
    np.testing.assert_allclose(left.detach().cpu().numpy(), g["left"], rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(right.detach().cpu().numpy(), g["right"], rtol=1e-3, atol=1e-6)
    # 1-D real-valued start, as get_samples builds it (gibbsSampler.py:57-61)
    start = np.concatenate([-166 * g["starts"][0] + 88, -166 * g["starts"][1] + 88]).astype(np.float32)
    gl, gr = gs.gibbs_sampling(torch.from_numpy(start).to(DEV), int(g["gk"]))
    assert gl.dim() == 1
    np.testing.assert_allclose(gl.detach().cpu().numpy(), g["gleft"], rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(gr.detach().cpu().numpy(), g["gright"], rtol=1e-3, atol=1e-6)


def test_meanfield_is_differentiable_like_reference():
    rng = np.random.default_rng(2)
    W = (rng.standard_normal((24, 16)) * 0.3).astype(np.float32)
    a = (0.1 * rng.standard_normal(24)).astype(np.float32); c = (0.1 * rng.standard_normal(16)).astype(np.float32)
    rbm = make_rbm(W, a, c)
    gs = divae_b200.GibbsSampler(RBM=rbm, n_gibbs_sampling_steps=2)
    x = torch.rand(5, 24, device=DEV, requires_grad=True)
    left, right = gs.gibbs_sampling(x, 2)
    (left.sum() + (right ** 2).sum()).backward()
    # torch reference of the same expression
    Wt = torch.tensor(W, device=DEV, requires_grad=True); at = torch.tensor(a, device=DEV, requires_grad=True)
    ct = torch.tensor(c, device=DEV, requires_grad=True); xt = x.detach().clone().requires_grad_(True)
    l = xt
    for _ in range(2):
        r = torch.sigmoid(l @ Wt + ct); l = torch.sigmoid(r @ Wt.t() + at)
    (l.sum() + (r ** 2).sum()).backward()
    for mine, ref in ((rbm._weights.grad, Wt.grad), (rbm._visible_bias.grad, at.grad),
                      (rbm._hidden_bias.grad, ct.grad), (x.grad, xt.grad)):
        torch.testing.assert_close(mine, ref, rtol=1e-3, atol=1e-5)


def test_energy_matches_reference_fixture(golden):
    g = golden("energy")
    rbm = make_rbm(g["W"], g["a"], g["c"])
    for nm in ("bin", "real"):
        z = torch.from_numpy(g["z" + nm[0]]).to(DEV)
        parts = [z[:, :30].contiguous(), z[:, 30:].contiguous()]
        np.testing.assert_allclose(rbm.energy(parts).detach().cpu().numpy(), g["E_" + nm], rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(rbm.log_prob(parts).detach().cpu().numpy(), g["logp_" + nm], rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(rbm.cross_entropy(parts).detach().cpu().numpy(), g["xent_" + nm], rtol=1e-3, atol=1e-4)


def test_energy_backward_matches_autograd_of_reference_expression(golden):
    g = golden("energy")
    rbm = make_rbm(g["W"], g["a"], g["c"])
    z = torch.from_numpy(g["zr"]).to(DEV).requires_grad_(True)
    rbm.energy([z]).mean().backward()
    Wt = torch.tensor(g["W"], device=DEV, requires_grad=True); at = torch.tensor(g["a"], device=DEV, requires_grad=True)
    ct = torch.tensor(g["c"], device=DEV, requires_grad=True); zt = z.detach().clone().requires_grad_(True)
    v, h = zt[:, :48], zt[:, 48:]
    (-(v @ at) - (h @ ct) - ((v @ Wt) * h).sum(1)).mean().backward()
    for mine, ref in ((rbm._weights.grad, Wt.grad), (rbm._visible_bias.grad, at.grad),
                      (rbm._hidden_bias.grad, ct.grad), (z.grad, zt.grad)):
        torch.testing.assert_close(mine, ref, rtol=1e-3, atol=1e-5)


def test_negative_phase_equals_reference_autograd(golden):
    g = golden("negphase")
    rbm = make_rbm(g["W"], g["a"], g["c"])
    v = torch.from_numpy(unpack(g["v"], 64)).to(DEV); h = torch.from_numpy(unpack(g["h"], 64)).to(DEV)
    S, sv, sh = _ops.negative_phase_stats(v, h)
    np.testing.assert_allclose(S.cpu().numpy(), g["gW"], atol=1e-6)
    np.testing.assert_allclose(sv.cpu().numpy(), g["ga"], atol=1e-6)
    np.testing.assert_allclose(sh.cpu().numpy(), g["gc"], atol=1e-6)
    e = _ops.energy_exp_from_stats(rbm, S, sv, sh)
    np.testing.assert_allclose(e.item(), g["e"], rtol=1e-4)
    (-e).backward()   # the reference backpropagates neg_energy = -energy_exp (gumbolt.py:104)
    np.testing.assert_allclose(rbm._weights.grad.cpu().numpy(), g["gW"], atol=1e-6)
    np.testing.assert_allclose(rbm._visible_bias.grad.cpu().numpy(), g["ga"], atol=1e-6)
    np.testing.assert_allclose(rbm._hidden_bias.grad.cpu().numpy(), g["gc"], atol=1e-6)


def test_negative_phase_large_batch_is_exact_counts():
    rng = np.random.default_rng(4)
    B, nV, nH = 5000, 70, 130
    v = (rng.random((B, nV)) < 0.3).astype(np.float32); h = (rng.random((B, nH)) < 0.6).astype(np.float32)
    S, sv, sh = _ops.negative_phase_stats(torch.from_numpy(v).to(DEV), torch.from_numpy(h).to(DEV), scale=1.0)
    assert (S.cpu().numpy() == v.T @ h).all() and (sv.cpu().numpy() == v.sum(0)).all() and (sh.cpu().numpy() == h.sum(0)).all()


def test_chain_sharding_is_invisible(golden):
    """Rows [off, off+n) of a big job == a shard launched with chain_offset=off."""
    g = golden("pcd_64x64")
    W, a, c = g["W"], g["a"], g["c"]
    rbm = make_rbm(W, a, c)
    full = divae_b200.PCD(batch_size=96, RBM=rbm, n_gibbs_sampling_steps=5, seed=3, variant="general")
    vf, hf = full.block_gibbs_sampling()
    for off, n in ((0, 48), (48, 48), (10, 7)):
        sh = divae_b200.PCD(batch_size=n, RBM=rbm, n_gibbs_sampling_steps=5, seed=3, variant="general", chain_offset=off)
        vs, hs = sh.block_gibbs_sampling()
        assert torch.equal(vs, vf[off:off + n]) and torch.equal(hs, hf[off:off + n])


def test_edge_cases_empty_and_single():
    rbm = make_rbm(*[x for x in (np.zeros((5, 3), np.float32), np.zeros(5, np.float32), np.zeros(3, np.float32))])
    s = divae_b200.PCD(batch_size=1, RBM=rbm, n_gibbs_sampling_steps=0, seed=0, variant="general")
    v, h = s.block_gibbs_sampling()      # k = 0: initial state only
    assert v.shape == (1, 5)
    s = divae_b200.PCD(batch_size=1, RBM=rbm, n_gibbs_sampling_steps=2, seed=0, variant="general")
    v, h = s.block_gibbs_sampling()
    assert v.shape == (1, 5) and h.shape == (1, 3)
    out = s.hidden_samples(torch.zeros(0, 5, device=DEV))
    assert out.shape == (0, 3)
    with pytest.raises(RuntimeError, match="size mismatch"):
        s.hidden_samples(torch.zeros(2, 4, device=DEV))


def test_marginals_within_3_sigma_of_exact_enumeration():
    rng = np.random.default_rng(2)
    W = orc.round_bf16((rng.standard_normal((6, 5)) * 1.0).astype(np.float32))
    a = (0.3 * rng.standard_normal(6)).astype(np.float32); c = (0.3 * rng.standard_normal(5)).astype(np.float32)
    ev, eh, evh = orc.exact_moments(W, a, c)
    B = 200000
    s = divae_b200.PCD(batch_size=B, RBM=make_rbm(W, a, c), n_gibbs_sampling_steps=30, seed=11, variant="general")
    v, h = s.block_gibbs_sampling()
    S, sv, sh = _ops.negative_phase_stats(v, h)
    for emp, ex in ((sv.cpu().numpy(), ev), (sh.cpu().numpy(), eh), (S.cpu().numpy(), evh)):
        sigma = np.sqrt(np.maximum(ex * (1 - ex), 1e-9) / B)
        assert (np.abs(emp - ex) < 3.5 * sigma + 2e-4).all()
