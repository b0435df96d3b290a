"""CD-1 on the vanilla 784x128 RBM (BASELINE config 1) against the fixture produced by the
reference's sandbox/rbm_standalone.py:75-121 with the same (broadcast) uniforms."""
import numpy as np
import pytest
import torch

from divae_b200.samplers.contrastiveDivergence import ContrastiveDivergence
from oracle import rbm_oracle as orc
from helpers import unpack

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_cd1_update_matches_reference_fixture(golden):
    g = golden("cd1")
    W, a, c = g["W"], g["a"], g["c"]
    cd = ContrastiveDivergence(784, 128, learning_rate=1e-3, momentum_coefficient=0.5, n_gibbs_sampling_steps=int(g["k"]),
                               weight_cost=1e-4, device=DEV, seed=1)
    cd._weights.copy_(torch.from_numpy(W)); cd._visible_bias.copy_(torch.from_numpy(a)); cd._hidden_bias.copy_(torch.from_numpy(c))
    v0 = torch.from_numpy(unpack(g["v0"], 784)).to(DEV)
    loss = cd.contrastive_divergence(v0, uniforms=torch.from_numpy(g["U"]).to(DEV))
    np.testing.assert_allclose(loss.item(), g["loss"], rtol=1e-4)
    np.testing.assert_allclose(cd._visible_bias.cpu().numpy(), g["a_new"], atol=2e-5)
    np.testing.assert_allclose(cd._hidden_bias.cpu().numpy(), g["c_new"], atol=2e-5)
    np.testing.assert_allclose(cd.weights_update.cpu().numpy()[::8, ::8], g["dW"], atol=5e-3, rtol=1e-3)
    np.testing.assert_allclose(cd._weights.cpu().numpy(), g["W_new"].astype(np.float32), atol=1e-3)


def test_cd1_statistics_match_oracle_with_independent_uniforms():
    rng = np.random.default_rng(0)
    nV, nH, B = 96, 40, 64
    W = (rng.standard_normal((nV, nH)) * 0.1).astype(np.float32)
    a = (0.1 * rng.standard_normal(nV)).astype(np.float32); c = (0.1 * rng.standard_normal(nH)).astype(np.float32)
    v0 = (rng.random((B, nV)) > 0.5).astype(np.float32)
    U = rng.random((3, nH)).astype(np.float32)
    cd = ContrastiveDivergence(nV, nH, 1e-3, 0.5, 2, 1e-4, device=DEV, seed=1)
    cd._weights.copy_(torch.from_numpy(W)); cd._visible_bias.copy_(torch.from_numpy(a)); cd._hidden_bias.copy_(torch.from_numpy(c))
    loss = cd.contrastive_divergence(torch.from_numpy(v0).to(DEV), uniforms=torch.from_numpy(U).to(DEV))
    st = orc.cd_k(v0, W, a, c, 2, lambda i, s: U[i][None, :])
    state = orc.cd_update(dict(W=W.copy(), a=a.copy(), c=c.copy(), dW=np.zeros_like(W), da=np.zeros_like(a), dc=np.zeros_like(c)), st)
    np.testing.assert_allclose(loss.item(), st["recon_sse"], rtol=1e-4)
    np.testing.assert_allclose(cd._weights.cpu().numpy(), state["W"], atol=1e-5)
    np.testing.assert_allclose(cd._visible_bias.cpu().numpy(), state["a"], atol=1e-5)
    np.testing.assert_allclose(cd._hidden_bias.cpu().numpy(), state["c"], atol=1e-5)


def test_cd_training_reduces_reconstruction_error_with_philox_draws():
    torch.manual_seed(0)
    data = (torch.rand(128, 784, device=DEV) < torch.linspace(0.05, 0.95, 784, device=DEV)).float()
    cd = ContrastiveDivergence(784, 128, 1e-3, 0.5, 1, 1e-4, device=DEV, seed=3)
    losses = [cd.contrastive_divergence(data).item() for _ in range(30)]
    assert losses[-1] < 0.8 * losses[0]


@pytest.mark.parametrize("B,k", [(128, 1), (200, 3), (1100, 2)])
def test_cd_k_fused_and_multi_launch_paths_match_oracle(B, k):
    """B <= 1024 runs as ONE cooperative launch (cd_fused_kernel), larger batches as 3k + 4 launches; both restate
    sandbox/rbm_standalone.py:75-121 and must agree with the oracle (ragged shape, broadcast uniforms)."""
    rng = np.random.default_rng(B + k)
    nV, nH = 203, 77
    W = (rng.standard_normal((nV, nH)) * 0.1).astype(np.float32)
    a = (0.1 * rng.standard_normal(nV)).astype(np.float32); c = (0.1 * rng.standard_normal(nH)).astype(np.float32)
    v0 = (rng.random((B, nV)) > 0.5).astype(np.float32)
    U = rng.random((k + 1, nH)).astype(np.float32)
    cd = ContrastiveDivergence(nV, nH, 1e-3, 0.5, k, 1e-4, device=DEV, seed=1)
    cd._weights.copy_(torch.from_numpy(W)); cd._visible_bias.copy_(torch.from_numpy(a)); cd._hidden_bias.copy_(torch.from_numpy(c))
    state = dict(W=W.copy(), a=a.copy(), c=c.copy(), dW=np.zeros_like(W), da=np.zeros_like(a), dc=np.zeros_like(c))
    for it in range(2):      # two updates: the second one starts from non-zero momentum buffers
        loss = cd.contrastive_divergence(torch.from_numpy(v0).to(DEV), uniforms=torch.from_numpy(U).to(DEV))
        st = orc.cd_k(v0, state["W"], state["a"], state["c"], k, lambda i, s: U[i][None, :])
        state = orc.cd_update(state, st)
        np.testing.assert_allclose(loss.item(), st["recon_sse"], rtol=2e-4)
    np.testing.assert_allclose(cd._weights.cpu().numpy(), state["W"], atol=2e-5)
    np.testing.assert_allclose(cd._visible_bias.cpu().numpy(), state["a"], atol=2e-5)
    np.testing.assert_allclose(cd._hidden_bias.cpu().numpy(), state["c"], atol=2e-5)
    np.testing.assert_allclose(cd.weights_update.cpu().numpy(), state["dW"], atol=2e-3, rtol=1e-4)
