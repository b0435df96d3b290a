"""GPU parity of the decoder path against the oracle and the reference-made fixtures.
Every call goes through the C ABI (rbm_b200_decoder_forward).

Tolerances: the layers run on bf16 tensor cores with operands split in high + low parts (three products,
fp32 accumulation), so outputs agree with the fp32 reference to ~2^-16 relative per layer; the tests ask for
|gpu - ref| <= 2e-4 * (1 + |ref|) after up to four layers with O(1..100) activations.  The hit gate is
bit-exact against the oracle's Philox gate except where |sigmoid(hits) - u| < 1e-4 (logit error budget)."""
import numpy as np
import pytest
import torch

import divae_b200
from divae_b200.decoder import DecoderRunner, generate_samples
from oracle import decoder_oracle as dec
from helpers import decoder_fixture_layers

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class V3(torch.nn.Module):
    """Attribute-for-attribute stand-in for the reference's BasicDecoderV3 (basicCoders.py:63-84, networks.py:58-80)."""

    def __init__(self, nodes, layers=None):
        super().__init__()
        self._activation_fct = torch.nn.ReLU()
        self._output_activation_fct = torch.nn.Identity()
        n = len(nodes)
        self._layers = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in nodes[:n - 2]])
        self._layers2 = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in nodes[n - 2:]])
        self._layers3 = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in nodes[n - 2:]])
        if layers is not None:
            for mods, ch in zip((self._layers, self._layers2, self._layers3), layers):
                for m, (W, b) in zip(mods, ch):
                    with torch.no_grad():
                        m.weight.copy_(torch.from_numpy(W)); m.bias.copy_(torch.from_numpy(b))

    def arrays(self):
        f = lambda mods: [(m.weight.detach().cpu().numpy(), m.bias.detach().cpu().numpy()) for m in mods]
        return f(self._layers), f(self._layers2), f(self._layers3)


class Basic(torch.nn.Module):
    """Stand-in for BasicDecoder (basicCoders.py:28-42)."""

    def __init__(self, nodes):
        super().__init__()
        self._activation_fct = torch.nn.ReLU()
        self._output_activation_fct = torch.nn.Identity()
        self._layers = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in nodes])


def close(gpu, ref, tol=2e-4):
    gpu = gpu.cpu().numpy().astype(np.float64); ref = np.asarray(ref, dtype=np.float64)
    err = np.abs(gpu - ref) / (1.0 + np.abs(ref))
    assert err.max() <= tol, err.max()


def ref64(chain, x, last_linear=True):
    x = np.asarray(x, dtype=np.float64)
    for i, (W, b) in enumerate(chain):
        x = x @ W.astype(np.float64).T + b.astype(np.float64)
        if not (last_linear and i == len(chain) - 1):
            x = np.maximum(x, 0)
    return x


def test_fixture_v3_forward_and_shower(golden):
    g = golden("decoder_v3")
    layers = decoder_fixture_layers(g)
    nodes = [(W.shape[1], W.shape[0]) for W, _ in layers[0]] + [(W.shape[1], W.shape[0]) for W, _ in layers[1]]
    m = V3(nodes, layers).to(DEV)
    r = DecoderRunner(m)
    x = torch.from_numpy(g["x"]).to(DEV)
    hits, act = r(x)
    close(hits, g["out_hits"]); close(act, g["out_act"])
    s = r.shower(x, seed=int(g["seed"]))
    _, clean = dec.shower_philox(g["out_hits"], g["out_act"], int(g["seed"]))
    sn = s.cpu().numpy()
    assert clean.mean() > 0.99
    # same gate as the reference's GumbelMod fed rho = 1 - u; activations within tolerance
    assert ((sn > 0) == (g["sample"] > 0))[clean & (g["out_act"] > 1e-3)].all()
    assert (sn[clean & (g["sample"] == 0)] == 0).all()
    close(torch.from_numpy(sn[clean]), g["sample"][clean])


def test_fixture_basic_decoder(golden):
    g = golden("decoder_basic")
    n = int(g["n_layers"])
    m = Basic([(g["W%d" % i].shape[1], g["W%d" % i].shape[0]) for i in range(n)])
    for i, lin in enumerate(m._layers):
        with torch.no_grad():
            lin.weight.copy_(torch.from_numpy(g["W%d" % i])); lin.bias.copy_(torch.from_numpy(g["b%d" % i]))
    y = DecoderRunner(m.to(DEV))(torch.from_numpy(g["x"]).to(DEV))
    close(y, g["y"])


@pytest.mark.parametrize("B", [300, 5000])
def test_calorimeter_shape_against_oracle(B):
    """gumboltcaloV.yaml: latent 2*64+1 -> 200 -> 300 -> [400 -> 504] x 2 (gumboltCaloV5.py:100-103)."""
    torch.manual_seed(11)
    m = V3([(129, 200), (200, 300), (300, 400), (400, 504)]).to(DEV)
    rng = np.random.default_rng(B)
    x = np.concatenate([(rng.random((B, 128)) < 0.5).astype(np.float32), (rng.random((B, 1)) * 100).astype(np.float32)], 1)
    trunk, hh, ha = m.arrays()
    r = DecoderRunner(m)
    xt = torch.from_numpy(x).to(DEV)
    hits, act = r(xt)
    t = ref64(trunk, x, last_linear=False)
    rh, ra = ref64(hh, t), ref64(ha, t)
    close(hits, rh); close(act, ra)
    # fp32 restatement agrees too (sanity of the tolerance)
    oh, oa = dec.decoder_v3(trunk, hh, ha, x)
    close(hits, oh, 5e-4); close(act, oa, 5e-4)
    # shower: gate bit-exact on clean entries, values relu(act) there
    seed, off, step = 99, 1000, 5
    s = r.shower(xt, seed=seed, chain_offset=off, step=step).cpu().numpy()
    so, clean = dec.shower_philox(rh, ra, seed, off, step)
    assert clean.mean() > 0.995
    assert ((s > 0) == (so > 0))[clean & (ra > 1e-3)].all()
    close(torch.from_numpy(s[clean]), so[clean])
    # hit rate follows sigmoid(hits)
    p = 1.0 / (1.0 + np.exp(-rh))
    gate = (s > 0)[ra > 1e-3].mean(); expect = p[ra > 1e-3].mean()
    assert abs(gate - expect) < 5 * np.sqrt(0.25 / (ra > 1e-3).sum()) + 1e-3


@pytest.mark.parametrize("nodes,B", [([(7, 10), (10, 70), (70, 130), (130, 33)], 37),
                                     ([(64, 64), (64, 64), (64, 64)], 128),
                                     ([(513, 260), (260, 1000), (1000, 300), (300, 777)], 1000)])
def test_ragged_shapes(nodes, B):
    torch.manual_seed(B)
    m = V3(nodes).to(DEV) if len(nodes) >= 3 else None
    rng = np.random.default_rng(B)
    x = rng.standard_normal((B, nodes[0][0])).astype(np.float32)     # real-valued inputs (posterior samples)
    trunk, hh, ha = m.arrays()
    hits, act = DecoderRunner(m)(torch.from_numpy(x).to(DEV))
    t = ref64(trunk, x, last_linear=False)
    close(hits, ref64(hh, t)); close(act, ref64(ha, t))
    b = Basic(nodes).to(DEV)
    chain = [(l.weight.detach().cpu().numpy(), l.bias.detach().cpu().numpy()) for l in b._layers]
    close(DecoderRunner(b)(torch.from_numpy(x).to(DEV)), ref64(chain, x))


def test_shards_see_the_same_draws():
    torch.manual_seed(2)
    m = V3([(33, 40), (40, 56), (56, 72), (72, 100)]).to(DEV)
    r = DecoderRunner(m)
    x = torch.randn(600, 33, device=DEV)
    full = r.shower(x, seed=5)
    part = r.shower(x[256:], seed=5, chain_offset=256)
    assert torch.equal(full[256:], part)
    assert not torch.equal(full, r.shower(x, seed=6))


def test_generate_samples_mirrors_the_reference_signature():
    """GumBoltCaloV5.generate_samples(num_samples, true_energy) (gumboltCaloV5.py:78-96) on a model stand-in."""
    torch.manual_seed(4)

    class Model:
        pass

    model = Model()
    rbm = divae_b200.RBM(64, 64).to(DEV)
    model.sampler = divae_b200.PCD(batch_size=128, RBM=rbm, n_gibbs_sampling_steps=10, seed=3)
    model.decoder = V3([(129, 200), (200, 300), (300, 400), (400, 504)]).to(DEV)
    e, s = generate_samples(model, num_samples=1000, true_energy=50.0, seed=8)
    assert e.shape == (896, 1) and s.shape == (896, 504) and (e == 50.0).all()
    assert (s >= 0).all() and 0.05 < (s > 0).float().mean() < 0.95
    e2, s2 = generate_samples(model, num_samples=10, seed=8)
    assert s2.shape == (128, 504) and (e2 >= 0).all() and (e2 <= 100).all()


def test_errors_are_loud():
    m = V3([(8, 8), (8, 8), (8, 8)])
    with pytest.raises(RuntimeError, match="CUDA-only"):
        DecoderRunner(m)(torch.zeros(4, 8))
    m = m.to(DEV)
    with pytest.raises(RuntimeError, match="size mismatch"):
        DecoderRunner(m)(torch.zeros(4, 9, device=DEV))
    m._activation_fct = torch.nn.Tanh()
    with pytest.raises(RuntimeError, match="ReLU"):
        DecoderRunner(m)
    with pytest.raises(RuntimeError, match="two-headed"):
        DecoderRunner(Basic([(8, 8)]).to(DEV)).shower(torch.zeros(4, 8, device=DEV), seed=1)


def test_staged_parameters_are_reused_and_refreshed():
    """The runner keeps the split parameters in its workspace (params_staged = 1) until a parameter changes."""
    torch.manual_seed(6)
    m = V3([(33, 40), (40, 56), (56, 72), (72, 100)]).to(DEV)
    r = DecoderRunner(m)
    x = torch.randn(300, 33, device=DEV)
    a1 = r.shower(x, seed=5)
    a2 = r.shower(x, seed=5)                       # second call: staged parameters
    assert torch.equal(a1, a2)
    h1, c1 = r(x[:100])                            # smaller batch, same workspace, still staged
    h2, c2 = DecoderRunner(m)(x[:100])
    assert torch.equal(h1, h2) and torch.equal(c1, c2)
    with torch.no_grad():
        m._layers[0].weight.mul_(1.5)              # in-place update bumps _version -> restaged
    b1 = r.shower(x, seed=5)
    assert not torch.equal(a1, b1)
    assert torch.equal(b1, DecoderRunner(m).shower(x, seed=5))
    big = r.shower(torch.randn(5000, 33, device=DEV), seed=5)   # larger batch: new workspace, restaged
    assert big.shape == (5000, 100)
