"""Seeded shape fuzzing of the tensor-core paths: random (ragged) priors, chain counts and decoder layouts.
Complements the hand-picked shapes of test_gpu_umma.py / test_gpu_decoder.py: tile clipping, split counts,
pair/single/multi-step selection and padding all depend on the shape."""
import numpy as np
import pytest
import torch

import divae_b200
from divae_b200.decoder import DecoderRunner
from oracle import rbm_oracle as orc

import os

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
# RBM_B200_FUZZ=<n> multiplies the number of cases and RBM_B200_FUZZ_SEED=<s> changes them (bug hunts; default: the fixed set)
FUZZ = max(1, int(os.environ.get("RBM_B200_FUZZ", "1")))
FUZZ_SEED = int(os.environ.get("RBM_B200_FUZZ_SEED", "0"))


def make_rbm(W, a, c):
    r = divae_b200.RBM(W.shape[0], W.shape[1])
    with torch.no_grad():
        r._weights.copy_(torch.from_numpy(W)); r._visible_bias.copy_(torch.from_numpy(a))
        r._hidden_bias.copy_(torch.from_numpy(c))
    return r.to(DEV)


def random_cases(n, seed):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        nV = int(rng.choice([rng.integers(1, 40), rng.integers(40, 300), rng.integers(300, 1100)]))
        nH = int(rng.choice([rng.integers(1, 40), rng.integers(40, 300), rng.integers(300, 1100)]))
        B = int(rng.choice([rng.integers(1, 130), rng.integers(130, 3000), rng.integers(3000, 30000)]))
        k = int(rng.integers(1, 4))
        yield nV, nH, B, k, int(rng.integers(1, 2 ** 31))


@pytest.mark.parametrize("case", list(random_cases(14 * FUZZ, 2024 + FUZZ_SEED)))
def test_umma_chains_and_statistics_on_random_shapes(case):
    nV, nH, B, k, seed = case
    rng = np.random.default_rng(seed)
    W = orc.round_bf16((rng.standard_normal((nV, nH)) * 0.4 / np.sqrt(max(nV, nH))).astype(np.float32))
    a = (0.3 * rng.standard_normal(nV)).astype(np.float32)
    c = (0.3 * rng.standard_normal(nH)).astype(np.float32)
    s = divae_b200.PCD(batch_size=B, RBM=make_rbm(W, a, c), n_gibbs_sampling_steps=k, seed=seed, variant="umma")
    v, h, (S, sv, sh), _e = s.block_gibbs_sampling(_stats_scale=1.0)
    # statistics are exact integer counts of the returned samples
    assert torch.equal(S.double(), v.double().t() @ h.double())
    assert torch.equal(sv.double(), v.double().sum(0)) and torch.equal(sh.double(), h.double().sum(0))
    # the first and the last chains follow the oracle (draws are keyed by the global chain id)
    n = min(B, 64)
    ov, oh, clean = orc.block_gibbs_philox(W, a, c, n, k, seed)
    vn, hn = v[:n].cpu().numpy(), h[:n].cpu().numpy()
    assert clean.mean() > 0.8
    assert (vn[clean] == ov[clean]).all() and (hn[clean] == oh[clean]).all()
    if B > n:
        ov, oh, clean = orc.block_gibbs_philox(W, a, c, n, k, seed, chain_offset=B - n)
        vn, hn = v[-n:].cpu().numpy(), h[-n:].cpu().numpy()
        assert (vn[clean] == ov[clean]).all() and (hn[clean] == oh[clean]).all()


class V3(torch.nn.Module):
    def __init__(self, nodes):
        super().__init__()
        self._activation_fct = torch.nn.ReLU()
        self._layers = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in nodes[:-2]])
        self._layers2 = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in nodes[-2:]])
        self._layers3 = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in nodes[-2:]])


def decoder_cases(n, seed):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        depth = int(rng.integers(3, 6))
        widths = [int(rng.choice([rng.integers(1, 70), rng.integers(70, 600)])) for _ in range(depth + 1)]
        yield widths, int(rng.choice([rng.integers(1, 300), rng.integers(300, 9000)])), int(rng.integers(1, 2 ** 31))


@pytest.mark.parametrize("case", list(decoder_cases(8 * FUZZ, 77 + FUZZ_SEED)))
def test_decoder_on_random_layouts(case):
    widths, B, seed = case
    torch.manual_seed(seed)
    nodes = list(zip(widths[:-1], widths[1:]))
    m = V3(nodes).to(DEV)
    x = torch.randn(B, widths[0], device=DEV)
    hits, act = DecoderRunner(m)(x)
    with torch.no_grad():
        t = x.double()
        for l in m._layers:
            t = torch.relu(t @ l.weight.double().t() + l.bias.double())
        outs = []
        for head in (m._layers2, m._layers3):
            y = t
            for i, l in enumerate(head):
                y = y @ l.weight.double().t() + l.bias.double()
                if i != len(head) - 1:
                    y = torch.relu(y)
            outs.append(y)
    for got, ref in zip((hits, act), outs):
        err = ((got.double() - ref).abs() / (1 + ref.abs())).max().item()
        assert err <= 2e-4, (widths, B, err)
    # shower: gate statistics follow sigmoid(hits); values are relu(act) where the gate fired
    s = DecoderRunner(m).shower(x, seed=seed)
    fired = s != 0
    assert torch.all(s >= 0)
    pos = outs[1] > 1e-3
    gate_rate = fired[pos].double().mean().item() if pos.any() else 0.0
    expect = torch.sigmoid(outs[0])[pos].mean().item() if pos.any() else 0.0
    n_eff = max(int(pos.sum().item()), 1)
    assert abs(gate_rate - expect) < 5 * np.sqrt(0.25 / n_eff) + 2e-3
    close = ((s.double() - torch.relu(outs[1])).abs() / (1 + outs[1].abs()))[fired]
    assert close.numel() == 0 or close.max().item() <= 2e-4


def small_cases(n, seed):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        nV, nH = int(rng.integers(1, 257)), int(rng.integers(1, 257))
        B = int(rng.choice([rng.integers(1, 40), rng.integers(40, 700), rng.integers(700, 9000)]))
        yield nV, nH, B, int(rng.integers(1, 6)), int(rng.integers(1, 2 ** 31)), str(rng.choice(["smem", "general", "auto"]))


@pytest.mark.parametrize("case", list(small_cases(12 * FUZZ, 9 + FUZZ_SEED)))
def test_small_priors_on_random_shapes(case):
    """Shared-memory / general / AUTO paths: chains against the oracle, negative-phase statistics exact."""
    nV, nH, B, k, seed, variant = case
    rng = np.random.default_rng(seed)
    W = orc.round_bf16((rng.standard_normal((nV, nH)) * 0.5 / np.sqrt(max(nV, nH))).astype(np.float32))
    a = (0.3 * rng.standard_normal(nV)).astype(np.float32)
    c = (0.3 * rng.standard_normal(nH)).astype(np.float32)
    s = divae_b200.PCD(batch_size=B, RBM=make_rbm(W, a, c), n_gibbs_sampling_steps=k, seed=seed, variant=variant)
    v, h, (S, sv, sh), _e = s.block_gibbs_sampling(_stats_scale=1.0)
    assert torch.equal(S.double(), v.double().t() @ h.double())
    assert torch.equal(sv.double(), v.double().sum(0)) and torch.equal(sh.double(), h.double().sum(0))
    n = min(B, 48)
    for off in sorted({0, B - n}):
        ov, oh, clean = orc.block_gibbs_philox(W, a, c, n, k, seed, chain_offset=off)
        vn, hn = v[off:off + n].cpu().numpy(), h[off:off + n].cpu().numpy()
        assert clean.mean() > 0.7
        assert (vn[clean] == ov[clean]).all() and (hn[clean] == oh[clean]).all(), (case, off)


def cd_cases(n, seed):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        yield (int(rng.integers(1, 400)), int(rng.integers(1, 300)), int(rng.choice([rng.integers(1, 40), rng.integers(40, 1024), rng.integers(1025, 1500)])),
               int(rng.integers(1, 4)), int(rng.integers(1, 2 ** 31)))


@pytest.mark.parametrize("case", list(cd_cases(6 * FUZZ, 5 + FUZZ_SEED)))
def test_cd_k_on_random_shapes(case):
    """CD-k (cooperative kernel for batches <= 1024, multi-launch path above) against the oracle's restatement of
    sandbox/rbm_standalone.py:75-121 with broadcast uniforms."""
    from divae_b200.samplers.contrastiveDivergence import ContrastiveDivergence
    nV, nH, B, k, seed = case
    rng = np.random.default_rng(seed)
    W = (rng.standard_normal((nV, nH)) * 0.1).astype(np.float32)
    a = (0.1 * rng.standard_normal(nV)).astype(np.float32); c = (0.1 * rng.standard_normal(nH)).astype(np.float32)
    v0 = (rng.random((B, nV)) > 0.5).astype(np.float32)
    U = rng.random((k + 1, nH)).astype(np.float32)
    cd = ContrastiveDivergence(nV, nH, 1e-3, 0.5, k, 1e-4, device=DEV, seed=1)
    cd._weights.copy_(torch.from_numpy(W)); cd._visible_bias.copy_(torch.from_numpy(a)); cd._hidden_bias.copy_(torch.from_numpy(c))
    loss = cd.contrastive_divergence(torch.from_numpy(v0).to(DEV), uniforms=torch.from_numpy(U).to(DEV))
    st = orc.cd_k(v0, W, a, c, k, lambda i, s: U[i][None, :])
    state = orc.cd_update(dict(W=W.copy(), a=a.copy(), c=c.copy(), dW=np.zeros_like(W), da=np.zeros_like(a), dc=np.zeros_like(c)), st)
    np.testing.assert_allclose(loss.item(), st["recon_sse"], rtol=5e-4)
    np.testing.assert_allclose(cd._weights.cpu().numpy(), state["W"], atol=5e-5)
    np.testing.assert_allclose(cd._visible_bias.cpu().numpy(), state["a"], atol=5e-5)
    np.testing.assert_allclose(cd._hidden_bias.cpu().numpy(), state["c"], atol=5e-5)
