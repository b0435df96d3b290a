"""Build-container only: re-run the real reference and compare with the committed fixtures.
Skipped wherever /root/reference is absent (e.g. the GPU box)."""
import numpy as np
import pytest

from oracle import _ref_shim
from helpers import unpack

pytestmark = pytest.mark.skipif(not _ref_shim.reference_available(), reason="reference tree not present")


def test_golden_pcd_regenerates_identically(golden, tmp_path, monkeypatch):
    from oracle import gen_golden
    _ref_shim.install()
    monkeypatch.setattr(gen_golden, "OUT", str(tmp_path))
    gen_golden.gen_pcd("pcd_64x64", 64, 64, 100, 10, 0.3, 0x5EED)
    new = dict(np.load(tmp_path / "pcd_64x64.npz"))
    old = golden("pcd_64x64")
    for key in ("W", "a", "c", "v", "h", "v0", "h1", "v1"):
        assert (new[key] == old[key]).all(), key


def test_torch_port_is_bit_identical_to_reference():
    import torch
    from oracle import gen_golden, ref_torch_port as port
    _ref_shim.install()
    from models.samplers.pcd import PCD
    W, a, c = gen_golden.make_params(32, 48, 0.3, 5)
    rbm = gen_golden.ref_rbm(W, a, c)
    torch.manual_seed(1)
    s = PCD(batch_size=17, RBM=rbm, n_gibbs_sampling_steps=4)
    with torch.no_grad():
        v, h = s.block_gibbs_sampling()
    torch.manual_seed(1)
    st = port.pcd_init_state(17, 32)
    with torch.no_grad():
        pv, ph, _ = port.block_gibbs_sampling(st, torch.from_numpy(W), torch.from_numpy(a), torch.from_numpy(c), 4)
    assert torch.equal(v, pv) and torch.equal(h, ph)


def test_reference_gumbolt_builds_on_top_of_the_dropin_classes():
    """Build container only: the REAL reference GumBolt (models/autoencoders/gumbolt.py, untouched) constructs
    its prior and sampler from this repo's classes once divae_b200.dropin.install() has run; forward
    (encoder/decoder, no sampler) runs; the KL term reaches our sampler (which refuses CPU tensors)."""
    import sys
    import types
    import torch
    _ref_shim.install()
    saved = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")}
    for k in saved:
        del sys.modules[k]
    try:
        import divae_b200
        from divae_b200 import dropin
        dropin.install()
        from models.autoencoders.gumbolt import GumBolt          # reference file
        ns = types.SimpleNamespace
        cfg = ns(model=ns(n_latent_hierarchy_lvls=2, n_latent_nodes=64, n_encoder_layer_nodes=200, n_encoder_layers=2,
                          decoder_hidden_nodes=[200, 200], encoder_hidden_nodes=[200, 200], beta_smoothing_fct=5, output_smoothing_fct=8,
                          activation_fct="relu", model_type="GumBolt"),
                 engine=ns(weight_decay_factor=1e-4, n_gibbs_sampling_steps=40), sampler=None)
        model = GumBolt(flat_input_size=[784], train_ds_mean=torch.full((784,), 0.3), activation_fct=torch.nn.ReLU(), cfg=cfg)
        model.create_networks()
        assert type(model.prior) is divae_b200.RBM and type(model.sampler) is divae_b200.PCD
        assert model.prior.weights.shape == (64, 64)                       # discreteVAE.py:121-122
        assert model.sampler.get_batch_size() == 128 and model.sampler.n_gibbs_sampling_steps == 40   # dvaepp.py:57
        keys = set(model.state_dict())
        assert {"prior._weights", "prior._visible_bias", "prior._hidden_bias", "sampler._RBM._weights"} <= keys
        x = (torch.rand(8, 784) < 0.3).float()
        out = model(x, is_training=True) if "is_training" in model.forward.__code__.co_varnames else model(x)
        assert out is not None
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            model.kl_divergence(out.post_logits, out.post_samples)        # gumbolt.py:69-107 -> sampler
    finally:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_decoder_oracle_matches_reference_at_the_calorimeter_shape():
    """gumboltcaloV.yaml shape: latent 2*64+1 -> 200 -> 300 -> [400 -> 504] x 2 heads (gumboltCaloV5.py:100-103)."""
    import torch
    from oracle import decoder_oracle as dec, gen_golden
    _ref_shim.install()
    from models.networks.basicCoders import BasicDecoderV3
    torch.manual_seed(3)
    nodes = [(129, 200), (200, 300), (300, 400), (400, 504)]
    d3 = BasicDecoderV3(node_sequence=nodes, activation_fct=torch.nn.ReLU(), cfg=None)
    rng = np.random.default_rng(5)
    x = np.concatenate([(rng.random((40, 128)) < 0.5).astype(np.float32), (rng.random((40, 1)) * 100).astype(np.float32)], 1)
    with torch.no_grad():
        hits, act = d3(torch.from_numpy(x))
    la = gen_golden._layer_arrays
    oh, oa = dec.decoder_v3(la(d3._layers), la(d3._layers2), la(d3._layers3), x)
    np.testing.assert_allclose(oh, hits.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(oa, act.numpy(), rtol=1e-4, atol=1e-4)


def test_decoder_fixture_regenerates_identically(golden, tmp_path, monkeypatch):
    from oracle import gen_golden
    _ref_shim.install()
    monkeypatch.setattr(gen_golden, "OUT", str(tmp_path))
    gen_golden.gen_decoder()
    for name in ("decoder_v3", "decoder_basic"):
        new, old = dict(np.load(tmp_path / (name + ".npz"))), golden(name)
        assert sorted(new) == sorted(old)
        for key in new:
            assert np.array_equal(new[key], old[key]), (name, key)


def test_encoder_fixture_regenerates_identically(golden, tmp_path, monkeypatch):
    from oracle import gen_golden
    _ref_shim.install()
    monkeypatch.setattr(gen_golden, "OUT", str(tmp_path))
    gen_golden.gen_encoder()
    new, old = dict(np.load(tmp_path / "encoder.npz")), golden("encoder")
    assert sorted(new) == sorted(old)
    for key in new:
        assert np.array_equal(new[key], old[key]), key
