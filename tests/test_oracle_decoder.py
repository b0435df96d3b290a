"""CPU: the decoder restatement (oracle/decoder_oracle.py) against the fixtures made by running the
reference's BasicDecoderV3 / BasicDecoder / GumbelMod (oracle/gen_golden.py:gen_decoder)."""
import numpy as np

from oracle import decoder_oracle as dec
from helpers import decoder_fixture_layers


def test_decoder_v3_and_shower_tail_match_reference(golden):
    g = golden("decoder_v3")
    trunk, hh, ha = decoder_fixture_layers(g)
    hits, act = dec.decoder_v3(trunk, hh, ha, g["x"])
    np.testing.assert_allclose(hits, g["out_hits"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(act, g["out_act"], rtol=1e-5, atol=1e-5)
    # the tail on the reference's own head outputs and the injected rho: bit-identical
    assert np.array_equal(dec.shower_tail(g["out_hits"], g["out_act"], g["rho"]), g["sample"])
    assert 0.2 < (g["sample"] > 0).mean() < 0.8          # a non-trivial gate


def test_philox_gate_equals_reference_gate_fed_rho_equals_one_minus_u(golden):
    g = golden("decoder_v3")
    s, clean = dec.shower_philox(g["out_hits"], g["out_act"], int(g["seed"]))
    assert clean.mean() > 0.99
    assert np.array_equal(s[clean], g["sample"][clean])
    u = dec.hit_uniforms(int(g["seed"]), *g["out_hits"].shape)
    np.testing.assert_array_equal((1.0 - u).astype(np.float32), g["rho"])


def test_decoder_basic_matches_reference(golden):
    g = golden("decoder_basic")
    layers = [(g["W%d" % i], g["b%d" % i]) for i in range(int(g["n_layers"]))]
    np.testing.assert_allclose(dec.decoder_basic(layers, g["x"]), g["y"], rtol=1e-5, atol=1e-5)


def test_hit_uniforms_follow_the_chain_offset():
    a = dec.hit_uniforms(7, 10, 40, chain_offset=0, step=3)
    b = dec.hit_uniforms(7, 4, 40, chain_offset=6, step=3)
    assert np.array_equal(a[6:], b)
    assert not np.array_equal(a, dec.hit_uniforms(7, 10, 40, chain_offset=0, step=4))


def _encoder_levels(g):
    return [[(g["L%d_W%d" % (l, i)], g["L%d_b%d" % (l, i)]) for i in range(int(g["n_layers"]))] for l in range(int(g["n_levels"]))]


def test_encoder_oracle_matches_reference_in_both_modes(golden):
    """tests/golden/encoder.npz: the reference's HierarchicalEncoder.forward (Gumbel smoother) fed the uniforms of
    the device contract; the restatement must reproduce logits and samples of every hierarchy level."""
    g = golden("encoder")
    levels, seed, beta = _encoder_levels(g), int(g["seed"]), float(g["beta"])
    n_lat = levels[0][-1][0].shape[0]
    for mode, is_training in (("train", True), ("eval", False)):
        pl, ps, clean = dec.encoder_forward_philox(levels, g["x"], beta, is_training, seed)
        for lvl in range(len(levels)):
            np.testing.assert_allclose(pl[lvl], g["%s_logits%d" % (mode, lvl)], rtol=1e-5, atol=1e-5)
            if is_training:
                np.testing.assert_allclose(ps[lvl], g["%s_samples%d" % (mode, lvl)], rtol=0, atol=2e-5)
            else:
                assert clean.mean() > 0.9
                assert np.array_equal(ps[lvl][clean], g["%s_samples%d" % (mode, lvl)][clean])
    # the injected-rho form is the same function
    rho = lambda lvl, shape: dec.smooth_rho(dec.smooth_u32(seed, shape[0], shape[1], 0, 0, lvl * n_lat))
    pl2, ps2 = dec.encoder_forward(levels, g["x"], beta, True, rho)
    np.testing.assert_allclose(ps2[1], g["train_samples1"], rtol=0, atol=2e-5)
