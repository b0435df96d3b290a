"""The oracle reproduces what the reference's own code returned (tests/golden, made by
oracle/gen_golden.py executing /root/reference), plus exact-enumeration known answers."""
import numpy as np
import pytest

from oracle import rbm_oracle as orc
from helpers import unpack

PCD_CASES = ["pcd_64x64", "pcd_256x256", "pcd_784x128", "pcd_200x200", "pcd_refinit_64x64"]


@pytest.mark.parametrize("name", PCD_CASES)
def test_block_gibbs_matches_reference(golden, name):
    g = golden(name)
    W, a, c = g["W"], g["a"], g["c"]
    B, k, seed = int(g["B"]), int(g["k"]), int(g["seed"])
    nV, nH = W.shape
    v0 = orc.init_state_philox(B, nV, seed)
    assert (v0 == unpack(g["v0"], nV)).all()
    v, h, clean = orc.block_gibbs_philox(W, a, c, B, k, seed)
    assert clean.mean() > 0.9
    assert (v[clean] == unpack(g["v"], nV)[clean]).all()
    assert (h[clean] == unpack(g["h"], nH)[clean]).all()
    # single half-steps (PCD.hidden_samples / visible_samples)
    draw = orc.philox_uniforms(seed, B)
    ph = orc.hidden_probs(v0, W, c)
    u = draw(0, ph.shape)
    ok = ~(np.abs(ph - u) < 1e-6).any(axis=1)
    h1 = orc.bernoulli(ph, u)
    assert (h1[ok] == unpack(g["h1"], nH)[ok]).all()


def test_energy_matches_reference(golden):
    g = golden("energy")
    for nm in ("bin", "real"):
        z = g["z" + nm[0]]
        E = orc.energy([z[:, :30], z[:, 30:]], g["W"], g["a"], g["c"])
        np.testing.assert_allclose(E, g["E_" + nm], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(orc.log_prob([z], g["W"], g["a"], g["c"]), g["logp_" + nm], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(-orc.log_prob([z], g["W"], g["a"], g["c"]), g["xent_" + nm], rtol=1e-5, atol=1e-5)
    assert abs(float(g["logZ"]) - orc.LOGZ_PLACEHOLDER) < 1e-5


def test_meanfield_matches_reference(golden):
    g = golden("meanfield")
    left, right = orc.meanfield_gibbs(g["x"], g["W"], g["a"], g["c"], int(g["k"]))
    np.testing.assert_allclose(left, g["left"], atol=2e-6)
    np.testing.assert_allclose(right, g["right"], atol=2e-6)
    start = np.concatenate([-166 * g["starts"][0] + 88, -166 * g["starts"][1] + 88]).astype(np.float32)
    gl, gr = orc.meanfield_gibbs(start[None, :], g["W"], g["a"], g["c"], int(g["gk"]))
    np.testing.assert_allclose(gl[0], g["gleft"], atol=2e-6)
    np.testing.assert_allclose(gr[0], g["gright"], atol=2e-6)


def test_cd1_matches_reference(golden):
    g = golden("cd1")
    W, a, c = g["W"], g["a"], g["c"]
    v0 = unpack(g["v0"], W.shape[0])
    st = orc.cd_k(v0, W, a, c, int(g["k"]), lambda i, s: g["U"][i][None, :])
    state = dict(W=W.copy(), a=a.copy(), c=c.copy(), dW=np.zeros_like(W), da=np.zeros_like(a), dc=np.zeros_like(c))
    state = orc.cd_update(state, st)
    np.testing.assert_allclose(state["a"], g["a_new"], atol=1e-5)
    np.testing.assert_allclose(state["c"], g["c_new"], atol=1e-5)
    np.testing.assert_allclose(state["dW"][::8, ::8], g["dW"], atol=2e-4)
    np.testing.assert_allclose(st["recon_sse"], g["loss"], rtol=1e-5)


def test_negative_phase_matches_reference_autograd(golden):
    g = golden("negphase")
    v, h = unpack(g["v"], 64), unpack(g["h"], 64)
    S, sv, sh = orc.negative_phase_stats(v, h)
    np.testing.assert_allclose(S, g["gW"], atol=1e-6)
    np.testing.assert_allclose(sv, g["ga"], atol=1e-6)
    np.testing.assert_allclose(sh, g["gc"], atol=1e-6)
    np.testing.assert_allclose(orc.energy_exp(v, h, g["W"], g["a"], g["c"]), g["e"], rtol=1e-5)


def test_port_check_fixture_is_binary(golden):
    g = golden("port_check")
    assert unpack(g["v"], 64).shape == (100, 64)


# ------------------------------------------------------------------ exact enumeration
def _small_rbm(nV, nH, scale, seed):
    rng = np.random.default_rng(seed)
    W = orc.round_bf16((rng.standard_normal((nV, nH)) * scale).astype(np.float32))
    a = (0.3 * rng.standard_normal(nV)).astype(np.float32)
    c = (0.3 * rng.standard_normal(nH)).astype(np.float32)
    return W, a, c


def test_exact_logZ_two_ways():
    W, a, c = _small_rbm(7, 9, 0.8, 1)
    z1 = orc.exact_logZ(W, a, c)
    z2 = orc.exact_logZ(W.T.copy(), c, a)  # swap roles of the layers
    assert abs(z1 - z2) < 1e-9
    # brute force over all joint states
    V = orc._all_states(7); H = orc._all_states(9)
    logp = (V @ a.astype(np.float64))[:, None] + (H @ c.astype(np.float64))[None, :] + V @ W.astype(np.float64) @ H.T
    assert abs(np.log(np.exp(logp).sum()) - z1) < 1e-9


def test_gibbs_chain_marginals_within_3_sigma_of_exact():
    W, a, c = _small_rbm(6, 5, 1.0, 2)
    ev, eh, evh = orc.exact_moments(W, a, c)
    B = 20000
    v, h, _ = orc.block_gibbs_philox(W, a, c, B, 30, seed=99)
    # returned (v, h): h ~ p(h | v_prev), v ~ p(v | h); both are (near-)stationary marginals
    for emp, ex in ((v.mean(0), ev), (h.mean(0), eh)):
        sigma = np.sqrt(ex * (1 - ex) / B)
        assert (np.abs(emp - ex) < 4 * sigma + 1e-3).all()
    S, _, _ = orc.negative_phase_stats(v, h)
    sigma = np.sqrt(np.maximum(evh * (1 - evh), 1e-6) / B)
    assert (np.abs(S - evh) < 4 * sigma + 1e-3).all()


def test_ais_oracle_close_to_exact_logZ():
    W, a, c = _small_rbm(12, 10, 0.5, 3)
    exact = orc.exact_logZ(W, a, c)
    est = orc.ais_logZ(W, a, c, n_chains=512, betas=np.linspace(0, 1, 501), seed=7)
    assert abs(est - exact) < 0.05


def test_bf16_rounding_is_idempotent_and_nearest():
    x = np.array([1.0, 1.00390625, 1.001953125, -3.1415926, 1e-3], dtype=np.float32)
    r = orc.round_bf16(x)
    assert (orc.round_bf16(r) == r).all()
    assert (np.abs(r - x) <= np.abs(x) * 2.0 ** -8).all()
    assert r[2] == 1.0  # tie rounds to even
