"""GPU parity of the posterior side: HierarchicalEncoder inference forward through
rbm_b200_decoder_forward (per-level MLP) + rbm_b200_gumbel_smooth (clamp, Gumbel smoothing / gate).
Tolerances as in test_gpu_decoder.py: logits within 2e-4 * (1 + |ref|); evaluation-mode samples bit-exact on rows
whose draws stay 1e-4 away from the threshold; training-mode samples within 2e-3 (beta = 7 amplifies logit error)."""
import numpy as np
import pytest
import torch

from divae_b200.decoder import EncoderRunner
from oracle import decoder_oracle as dec

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class Encoder(torch.nn.Module):
    """Attribute-for-attribute stand-in for the reference's HierarchicalEncoder (hierarchicalEncoder.py:25-98)."""

    def __init__(self, n_in, n_latent, levels, hidden0, hidden, beta):
        super().__init__()
        from types import SimpleNamespace
        self._config = SimpleNamespace(model=SimpleNamespace(beta_smoothing_fct=beta))
        self.smoothing_dist_mod = None
        nets = []
        for lvl in range(levels):
            sizes = ([n_in] + hidden0 + [n_latent]) if lvl == 0 else ([n_in + lvl * n_latent] + hidden + [n_latent])
            mods = []
            for i in range(len(sizes) - 1):
                mods.append(torch.nn.Linear(sizes[i], sizes[i + 1]))
                mods.append(torch.nn.Identity() if i == len(sizes) - 2 else torch.nn.ReLU())
            nets.append(torch.nn.Sequential(*mods))
        self._networks = torch.nn.ModuleList(nets)

    def arrays(self):
        return [[(m.weight.detach().cpu().numpy(), m.bias.detach().cpu().numpy()) for m in net if isinstance(m, torch.nn.Linear)]
                for net in self._networks]


def load_fixture(g):
    n_levels, n_layers = int(g["n_levels"]), int(g["n_layers"])
    W = [[g["L%d_W%d" % (l, i)] for i in range(n_layers)] for l in range(n_levels)]
    enc = Encoder(W[0][0].shape[1], W[0][-1].shape[0], n_levels, [w.shape[0] for w in W[0][:-1]], [w.shape[0] for w in W[1][:-1]],
                  float(g["beta"]))
    for l, net in enumerate(enc._networks):
        lin = [m for m in net if isinstance(m, torch.nn.Linear)]
        for i, m in enumerate(lin):
            with torch.no_grad():
                m.weight.copy_(torch.from_numpy(g["L%d_W%d" % (l, i)])); m.bias.copy_(torch.from_numpy(g["L%d_b%d" % (l, i)]))
    return enc.to(DEV)


def close(gpu, ref, tol):
    err = np.abs(gpu.detach().cpu().numpy().astype(np.float64) - ref) / (1.0 + np.abs(ref))
    assert err.max() <= tol, err.max()


def test_fixture_from_the_reference_encoder(golden):
    g = golden("encoder")
    r = EncoderRunner(load_fixture(g))
    x = torch.from_numpy(g["x"]).to(DEV)
    beta, logits, samples = r(x, is_training=True, seed=int(g["seed"]))
    assert float(beta) == float(g["beta"])
    for lvl in range(2):
        close(logits[lvl], g["train_logits%d" % lvl], 2e-4)
        close(samples[lvl], g["train_samples%d" % lvl], 2e-3)
    _, logits, samples = r(x, is_training=False, seed=int(g["seed"]))
    levels = [[(g["L%d_W%d" % (l, i)], g["L%d_b%d" % (l, i)]) for i in range(int(g["n_layers"]))] for l in range(2)]
    _, _, clean = dec.encoder_forward_philox(levels, g["x"], float(g["beta"]), False, int(g["seed"]))
    assert clean.mean() > 0.9
    for lvl in range(2):
        assert np.array_equal(samples[lvl].cpu().numpy()[clean], g["eval_samples%d" % lvl][clean])
        close(logits[lvl][torch.from_numpy(clean).to(DEV)], g["eval_logits%d" % lvl][clean], 2e-4)


@pytest.mark.parametrize("B", [100, 4000])
def test_calorimeter_encoder_shape(B):
    """gumboltcaloV.yaml: 504 voxels -> [400, 300, 200] -> 64 latents, second level (504 + 64) -> [64, 64] -> 64."""
    torch.manual_seed(B)
    enc = Encoder(504, 64, 2, [400, 300, 200], [64, 64], 7.0).to(DEV)
    rng = np.random.default_rng(B)
    x = np.maximum(rng.standard_normal((B, 504)), 0).astype(np.float32)      # sparse non-negative showers
    r = EncoderRunner(enc)
    xt = torch.from_numpy(x).to(DEV)
    seed, off, step = 17, 300, 2
    _, logits, samples = r(xt, is_training=False, seed=seed, chain_offset=off, step=step)
    pl, ps, clean = dec.encoder_forward_philox(enc.arrays(), x, 7.0, False, seed, off, step)
    assert clean.mean() > 0.9
    cm = torch.from_numpy(clean).to(DEV)
    for lvl in range(2):
        assert set(np.unique(samples[lvl].cpu().numpy())) <= {0.0, 1.0}
        assert np.array_equal(samples[lvl].cpu().numpy()[clean], ps[lvl][clean])
        close(logits[lvl][cm], pl[lvl][clean], 3e-4)
    _, logits, samples = r(xt, is_training=True, seed=seed, chain_offset=off, step=step)
    pl, ps, _ = dec.encoder_forward_philox(enc.arrays(), x, 7.0, True, seed, off, step)
    for lvl in range(2):
        close(logits[lvl], pl[lvl], 1e-3)          # level 1 sees level 0's real-valued samples: errors compound
        close(samples[lvl], ps[lvl], 5e-3)
        assert 0.0 <= samples[lvl].min().item() and samples[lvl].max().item() <= 1.0
        assert samples[lvl].requires_grad        # training mode with trainable parameters: the differentiable path


def test_clamp_sharding_and_errors():
    torch.manual_seed(0)
    enc = Encoder(20, 8, 2, [16], [12], 5.0).to(DEV)
    with torch.no_grad():
        enc._networks[0][-2].bias.fill_(500.0)        # logits beyond +88 must come back clamped
    r = EncoderRunner(enc)
    x = torch.randn(300, 20, device=DEV)
    _, logits, samples = r(x, is_training=False, seed=3)
    assert logits[0].max().item() == 88.0 and (samples[0] == 1).all()
    _, _, part = r(x[256:], is_training=False, seed=3, chain_offset=256)
    assert torch.equal(samples[1][256:], part[1])
    _, _, scaled = r((x, torch.full((300, 1), 2.0, device=DEV)), is_training=False, seed=3)
    assert torch.equal(scaled[0], 2.0 * samples[0])
    with pytest.raises(RuntimeError, match="CUDA-only"):
        EncoderRunner(Encoder(20, 8, 2, [16], [12], 5.0))(torch.zeros(4, 20))
