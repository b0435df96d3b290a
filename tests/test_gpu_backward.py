"""Backward passes of the rows next to the path on the tensor cores (SURVEY.md section 8f ranks 1 and 4,
row a11).  The reference gets these gradients from torch autograd over ATen matmuls:
  energy_exp on posterior samples          models/autoencoders/gumbolt.py:97-99,118-132
  mean-field half-steps                    models/samplers/gibbsSampler.py:20-28
  hierarchical encoder + Gumbel smoother   models/networks/hierarchicalEncoder.py:139-183, utils/dists/gumbelmod.py:14-24
Here every product is rbm_b200_gemm / rbm_b200_gemm_tn (tcgen05, split-bf16 operands).  Parity: against torch.autograd of the
reference expression evaluated in float64 on the same inputs, 1e-3 relative (north star: fp32 accumulate).
"""
import numpy as np
import pytest
import torch

import divae_b200
from divae_b200 import _ops
from divae_b200.decoder import EncoderRunner
from oracle import decoder_oracle as dec
from test_gpu_encoder import Encoder

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel_close(got, ref, tol=1e-3):
    got = got.double(); ref = ref.double()
    scale = ref.abs().max().clamp_min(1e-12)
    err = (got - ref).abs().max() / scale
    assert err <= tol, float(err)


@pytest.mark.parametrize("M,N,K", [(130, 200, 70), (1000, 504, 400), (64, 64, 64), (4097, 1024, 1024), (5, 3, 7), (300, 129, 568)])
@pytest.mark.parametrize("trans_b", [True, False])
def test_gemm_matches_float64(M, N, K, trans_b):
    g = torch.Generator(device="cpu").manual_seed(M + 3 * N + 7 * K)
    a = torch.randn(M, K, generator=g).to(DEV)
    b = (torch.randn(N, K, generator=g) if trans_b else torch.randn(K, N, generator=g)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    ref = a.double() @ (b.double().t() if trans_b else b.double())
    bound = a.double().abs() @ (b.double().abs().t() if trans_b else b.double().abs())   # |a| |b|: the natural error scale
    got = _ops.gemm(a, b, trans_b)
    assert got.shape == (M, N) and got.dtype == torch.float32
    assert ((got.double() - ref).abs() <= 6e-5 * bound + 1e-6).all(), float(((got.double() - ref).abs() / (bound + 1e-9)).max())
    got2 = _ops.gemm(a, b, trans_b, bias=bias, relu=True)
    ref2 = torch.relu(ref + bias.double())
    assert ((got2.double() - ref2).abs() <= 6e-5 * (bound + bias.double().abs()) + 1e-6).all()


@pytest.mark.parametrize("R,M,N", [(128, 64, 64), (1000, 200, 130), (4099, 504, 400), (7, 5, 3), (65536, 256, 256)])
def test_gemm_tn_matches_float64(R, M, N):
    g = torch.Generator(device="cpu").manual_seed(R + 3 * M + 7 * N)
    a = torch.randn(R, M, generator=g).to(DEV)
    b = torch.randn(R, N, generator=g).to(DEV)
    ref = a.double().t() @ b.double()
    bound = a.double().abs().t() @ b.double().abs()
    got = _ops.gemm_tn(a, b)
    assert got.shape == (M, N)
    assert ((got.double() - ref).abs() <= 6e-5 * bound + 1e-6).all(), float(((got.double() - ref).abs() / (bound + 1e-9)).max())


def make_rbm(nV, nH, scale, seed):
    torch.manual_seed(seed)
    r = divae_b200.RBM(nV, nH)
    with torch.no_grad():
        r._weights.copy_(torch.randn(nV, nH) * scale); r._visible_bias.copy_(torch.randn(nV) * 0.3)
        r._hidden_bias.copy_(torch.randn(nH) * 0.3)
    return r.to(DEV)


@pytest.mark.parametrize("B,nV,nH", [(128, 64, 64), (1024, 256, 256), (500, 200, 130)])
def test_energy_exp_backward_matches_autograd_of_the_reference_expression(B, nV, nH):
    """gumbolt.py:118-132: mean_b(-v^T W h - a.v - c.h) on smoothed posterior samples, gradients w.r.t. the samples
    (they flow back into the encoder) and (W, a, c)."""
    rbm = make_rbm(nV, nH, 0.2, 3)
    g = torch.Generator(device="cpu").manual_seed(5)
    zv = torch.rand(B, nV, generator=g).to(DEV).requires_grad_(True)
    zh = torch.rand(B, nH, generator=g).to(DEV).requires_grad_(True)
    e = divae_b200.energy_exp(rbm, zv, zh)
    e.backward()
    got = [zv.grad, zh.grad, rbm._weights.grad, rbm._visible_bias.grad, rbm._hidden_bias.grad]
    # the reference expression in float64
    W, a, c = (p.detach().double().requires_grad_(True) for p in (rbm._weights, rbm._visible_bias, rbm._hidden_bias))
    v64 = zv.detach().double().requires_grad_(True); h64 = zh.detach().double().requires_grad_(True)
    w = W + torch.zeros((B,) + W.size(), device=DEV, dtype=torch.float64)
    batch_energy = (-torch.matmul(v64.unsqueeze(2).permute(0, 2, 1), torch.matmul(w, h64.unsqueeze(2))).reshape(-1)
                    - torch.matmul(v64, a) - torch.matmul(h64, c))
    ref_e = torch.mean(batch_energy, 0)
    ref_e.backward()
    assert abs(e.item() - ref_e.item()) <= 1e-4 * max(1.0, abs(ref_e.item()))
    for x, y in zip(got, [v64.grad, h64.grad, W.grad, a.grad, c.grad]):
        rel_close(x, y)


@pytest.mark.parametrize("B,nV,nH,k", [(64, 64, 64, 3), (300, 200, 130, 2)])
def test_mean_field_backward_matches_autograd(B, nV, nH, k):
    """GibbsSampler.gibbs_sampling (gibbsSampler.py:31-37) is differentiable in the reference."""
    rbm = make_rbm(nV, nH, 0.1, 7)
    s = divae_b200.GibbsSampler(RBM=rbm, n_gibbs_sampling_steps=k)
    g = torch.Generator(device="cpu").manual_seed(9)
    x = torch.rand(B, nV, generator=g).to(DEV).requires_grad_(True)
    wl = torch.randn(B, nV, generator=g).to(DEV); wr = torch.randn(B, nH, generator=g).to(DEV)
    left, right = s.gibbs_sampling(x, k)
    ((left * wl).sum() + (right * wr).sum()).backward()
    got = [x.grad, rbm._weights.grad, rbm._visible_bias.grad, rbm._hidden_bias.grad]
    W, a, c = (p.detach().double().requires_grad_(True) for p in (rbm._weights, rbm._visible_bias, rbm._hidden_bias))
    x64 = x.detach().double().requires_grad_(True)
    l64 = x64
    for _ in range(k):
        r64 = torch.sigmoid(torch.matmul(l64, W) + c)          # gibbsSampler.py:20-23
        l64 = torch.sigmoid(torch.matmul(r64, W.t()) + a)      # gibbsSampler.py:25-28
    ((l64 * wl.double()).sum() + (r64 * wr.double()).sum()).backward()
    for u, v in zip(got, [x64.grad, W.grad, a.grad, c.grad]):
        rel_close(u, v)


def test_encoder_training_step_backward_matches_autograd():
    """HierarchicalEncoder.forward in training mode (hierarchicalEncoder.py:139-183) with the Gumbel smoother: the loss
    touches logits AND smoothed samples of both levels; gradients w.r.t. the input and every Linear parameter."""
    torch.manual_seed(11)
    beta, B, n_in, n_lat = 5.0, 200, 504, 64
    enc = Encoder(n_in, n_lat, 2, [400, 300, 200], [64, 64], beta).to(DEV)
    runner = EncoderRunner(enc)
    g = torch.Generator(device="cpu").manual_seed(13)
    x = torch.rand(B, n_in, generator=g).to(DEV).requires_grad_(True)
    wz = [torch.randn(B, n_lat, generator=g).to(DEV) for _ in range(2)]
    wl = [torch.randn(B, n_lat, generator=g).to(DEV) * 0.1 for _ in range(2)]
    _, logits, samples = runner(x, is_training=True, seed=21, step=3)
    assert samples[0].requires_grad and logits[1].requires_grad
    sum((z * w).sum() + (l * v).sum() for z, w, l, v in zip(samples, wz, logits, wl)).backward()
    got_x = x.grad.clone()
    got_p = [p.grad.clone() for p in enc.parameters()]
    # float64 torch graph of the same computation with the device's own rho (oracle restatement of the draw contract)
    rho = [torch.from_numpy(dec.smooth_rho(dec.smooth_u32(21, B, n_lat, 0, 3, lvl * n_lat)).astype(np.float64)).to(DEV) for lvl in range(2)]
    enc64 = Encoder(n_in, n_lat, 2, [400, 300, 200], [64, 64], beta).to(DEV).double()
    enc64.load_state_dict({k: v.double() for k, v in enc.state_dict().items()})
    x64 = x.detach().double().requires_grad_(True)
    post, lg = [], []
    for lvl, net in enumerate(enc64._networks):
        l = torch.clamp(net(torch.cat([x64] + post, dim=1)), min=-88., max=88.)
        z = torch.sigmoid((l + torch.log(rho[lvl]) - torch.log(1 - rho[lvl])) * beta)
        lg.append(l); post.append(z)
    sum((z * w.double()).sum() + (l * v.double()).sum() for z, w, l, v in zip(post, wz, lg, wl)).backward()
    for z, z64 in zip(samples, post):
        assert (z.double() - z64).abs().max() < 2e-3
    rel_close(got_x, x64.grad, 2e-3)
    for gp, p64 in zip(got_p, enc64.parameters()):
        rel_close(gp, p64.grad, 2e-3)
