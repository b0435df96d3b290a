"""ABI coverage and behavioural checks that are not tied to one kernel variant: the C entry
rbm_b200_meanfield, CUDA-graph capture, stream semantics, edge cases, and a caller-level
check that mimics how the reference's KL term consumes the sampler (gumbolt.py:97-106)."""
import ctypes

import numpy as np
import pytest
import torch

import divae_b200
from divae_b200 import _lib, _ops
from oracle import rbm_oracle as orc
from test_gpu_umma import make_rbm, rand_params, run

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_c_entry_meanfield_matches_oracle():
    lib = _lib.load()
    W, a, c = rand_params(70, 50, 0.2, 1)
    rbm = make_rbm(W, a, c)
    p, keep = rbm._pack.pack(rbm, need_bf16=False)
    x = np.random.default_rng(0).random((9, 70)).astype(np.float32)
    left = torch.from_numpy(x).to(DEV)
    right = torch.empty(9, 50, device=DEV)
    _lib.check(lib.rbm_b200_meanfield(ctypes.byref(p), _ops._ptr(left), _ops._ptr(right), 9, 6, _ops._stream()))
    ol, orr = orc.meanfield_gibbs(x, W, a, c, 6)
    np.testing.assert_allclose(left.cpu().numpy(), ol, rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(right.cpu().numpy(), orr, rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("variant,n,B", [("smem", 64, 100), ("umma", 320, 300), ("general", 40, 33)])
def test_block_gibbs_is_cuda_graph_capturable_and_stream_ordered(variant, n, B):
    """The library never synchronises nor touches the default stream: a call can be captured
    into a CUDA graph on a side stream and replayed."""
    lib = _lib.load()
    W, a, c = rand_params(n, n, 0.2, 2)
    rbm = make_rbm(W, a, c)
    p, keep = rbm._pack.pack(rbm, need_bf16=True)
    k = 3
    v = torch.zeros(B, n, device=DEV); h = torch.zeros(B, n, device=DEV)
    nbytes = lib.rbm_b200_block_gibbs_workspace(n, n, B, _lib.VARIANTS[variant])
    ws = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=DEV)
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        # warm-up outside capture (function attributes, driver entry point lookup)
        _lib.check(lib.rbm_b200_block_gibbs(ctypes.byref(p), _ops._ptr(v), _ops._ptr(h), B, k, 1, 5, 0, 0,
                                            _lib.VARIANTS[variant], _ops._ptr(ws), nbytes, ctypes.c_void_p(side.cuda_stream)))
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            _lib.check(lib.rbm_b200_block_gibbs(ctypes.byref(p), _ops._ptr(v), _ops._ptr(h), B, k, 1, 5, 0, 0,
                                                _lib.VARIANTS[variant], _ops._ptr(ws), nbytes,
                                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    v.zero_(); h.zero_()
    g.replay()
    torch.cuda.synchronize()
    ov, oh, clean = orc.block_gibbs_philox(W, a, c, B, k, 5)
    assert (v.cpu().numpy()[clean] == ov[clean]).all() and (h.cpu().numpy()[clean] == oh[clean]).all()


@pytest.mark.parametrize("variant", ["smem", "umma", "general"])
def test_edge_cases_k0_single_chain_and_missing_workspace(variant):
    W, a, c = rand_params(48, 48, 0.2, 3)
    v, h = run(W, a, c, 1, 0, seed=1, variant=variant)          # k = 0: the initial state comes back
    assert (v == orc.init_state_philox(1, 48, 1)).all()
    v, h = run(W, a, c, 1, 2, seed=1, variant=variant)
    ov, oh, clean = orc.block_gibbs_philox(W, a, c, 1, 2, 1)
    assert clean.all() and (v == ov).all() and (h == oh).all()
    if variant == "umma":
        lib = _lib.load()
        rbm = make_rbm(W, a, c)
        p, keep = rbm._pack.pack(rbm, need_bf16=True)
        x = torch.zeros(4, 48, device=DEV)
        rc = lib.rbm_b200_block_gibbs(ctypes.byref(p), _ops._ptr(x), _ops._ptr(x), 4, 1, 1, 0, 0, 0, _lib.VARIANT_UMMA,
                                      None, 0, _ops._stream())
        assert rc == _lib.ERR_WORKSPACE and b"workspace" in lib.rbm_b200_last_error()


def test_smem_variant_rejects_large_priors_and_umma_needs_workspace():
    lib = _lib.load()
    W, a, c = rand_params(300, 64, 0.1, 4)
    rbm = make_rbm(W, a, c)
    p, keep = rbm._pack.pack(rbm, need_bf16=True)
    x = torch.zeros(4, 300, device=DEV); y = torch.zeros(4, 64, device=DEV)
    rc = lib.rbm_b200_block_gibbs(ctypes.byref(p), _ops._ptr(x), _ops._ptr(y), 4, 1, 1, 0, 0, 0, _lib.VARIANT_SMEM, None, 0, _ops._stream())
    assert rc == _lib.ERR_UNSUPPORTED
    p2, keep2 = rbm._pack.pack(rbm, need_bf16=False)
    rc = lib.rbm_b200_block_gibbs(ctypes.byref(p2), _ops._ptr(x), _ops._ptr(y), 4, 1, 1, 0, 0, 0, _lib.VARIANT_UMMA, None, 0, _ops._stream())
    # W_bf16 is optional (the library rounds the fp32 W itself): what is missing here is the workspace
    assert rc == _lib.ERR_WORKSPACE and b"workspace" in lib.rbm_b200_last_error()


def test_bf16_weight_cache_follows_parameter_updates():
    W, a, c = rand_params(64, 64, 0.3, 5)
    rbm = make_rbm(W, a, c)
    s = divae_b200.PCD(batch_size=64, RBM=rbm, n_gibbs_sampling_steps=4, seed=3, variant="smem")
    s.block_gibbs_sampling()
    W2 = orc.round_bf16(W * 0.5)
    with torch.no_grad():
        rbm._weights.copy_(torch.from_numpy(W2).to(DEV))       # an optimiser step bumps W._version
    s2 = divae_b200.PCD(batch_size=64, RBM=rbm, n_gibbs_sampling_steps=4, seed=3, variant="smem")
    s2._pack = s._pack                                          # reuse the cache that saw the old W
    v, h = s2.block_gibbs_sampling()
    ov, oh, clean = orc.block_gibbs_philox(W2, a, c, 64, 4, 3)
    assert (v.cpu().numpy()[clean] == ov[clean]).all()


def test_kl_negative_phase_as_the_reference_consumes_it():
    """gumbolt.py:102-106: samples -> detach -> energy_exp -> backward.  Both the reference expression
    (broadcast W to [B,nV,nH]) on our samples and the fused negative_phase() give the same loss/grads."""
    W, a, c = rand_params(64, 64, 0.3, 6)
    rbm = make_rbm(W, a, c)
    sampler = divae_b200.PCD(batch_size=128, RBM=rbm, n_gibbs_sampling_steps=40, seed=11)
    rbm_visible_samples, rbm_hidden_samples = sampler.block_gibbs_sampling()
    rbm_vis, rbm_hid = rbm_visible_samples.detach(), rbm_hidden_samples.detach()
    assert not rbm_visible_samples.requires_grad and rbm_vis.dtype == torch.float32
    # the reference's energy_exp, verbatim arithmetic (gumbolt.py:118-134)
    w = rbm.weights + torch.zeros((rbm_vis.size(0),) + rbm.weights.size(), device=rbm_vis.device)
    vis = rbm_vis.unsqueeze(2).permute(0, 2, 1); hid = rbm_hid.unsqueeze(2)
    batch_energy = (-torch.matmul(vis, torch.matmul(w, hid)).reshape(-1) - torch.matmul(rbm_vis, rbm.visible_bias)
                    - torch.matmul(rbm_hid, rbm.hidden_bias))
    neg_energy_ref = -torch.mean(batch_energy, 0)
    gW, ga, gc = torch.autograd.grad(neg_energy_ref, [rbm._weights, rbm._visible_bias, rbm._hidden_bias])
    S, sv, sh = _ops.negative_phase_stats(rbm_vis, rbm_hid)
    e = _ops.energy_exp_from_stats(rbm, S, sv, sh)
    torch.testing.assert_close(-e, neg_energy_ref, rtol=1e-4, atol=1e-4)
    g2 = torch.autograd.grad(-e, [rbm._weights, rbm._visible_bias, rbm._hidden_bias])
    for mine, ref in zip(g2, (gW, ga, gc)):
        torch.testing.assert_close(mine, ref, rtol=1e-4, atol=1e-5)


def test_pcd_training_moves_model_marginals_toward_data():
    """A tiny RBM trained with the fused negative phase (positive phase from data, as in CD/PCD)
    learns the data's visible marginals: functional check of sampler + statistics + gradients."""
    torch.manual_seed(0)
    nV, nH, B = 32, 16, 512
    target = torch.linspace(0.1, 0.9, nV, device=DEV)
    rbm = divae_b200.RBM(nV, nH).to(DEV)
    sampler = divae_b200.PCD(batch_size=B, RBM=rbm, n_gibbs_sampling_steps=5, seed=1, persistent=True)
    opt = torch.optim.Adam(rbm.parameters(), lr=0.05)
    for it in range(150):
        data = (torch.rand(B, nV, device=DEV) < target).float()
        ph = sampler.hidden_samples(data)                                    # positive phase
        pos = rbm.energy([data, ph]).mean() if nV == nH else None
        Sp, svp, shp = _ops.negative_phase_stats(data, ph)
        e_pos = _ops.energy_exp_from_stats(rbm, Sp, svp, shp)
        e_neg, _ = sampler.negative_phase()
        loss = e_pos - e_neg                                                  # d/dtheta = -<.>_data + <.>_model
        opt.zero_grad(); loss.backward(); opt.step()
    probe = divae_b200.PCD(batch_size=8192, RBM=rbm, n_gibbs_sampling_steps=50, seed=2)
    v, _ = probe.block_gibbs_sampling()
    assert (v.mean(0) - target).abs().max().item() < 0.08


def test_energy_exp_on_posterior_samples_matches_reference_expression():
    """the positive-phase energy term on smoothed (real-valued) posterior samples,
    value and gradients w.r.t. samples and parameters vs the reference's broadcast expression."""
    W, a, c = rand_params(64, 64, 0.3, 21)
    rbm = make_rbm(W, a, c)
    torch.manual_seed(1)
    z_vis = torch.rand(100, 64, device=DEV, requires_grad=True)
    z_hid = torch.rand(100, 64, device=DEV, requires_grad=True)
    e = divae_b200.energy_exp(rbm, z_vis, z_hid)
    g = torch.autograd.grad(e, [z_vis, z_hid, rbm._weights, rbm._visible_bias, rbm._hidden_bias])
    # gumbolt.py:118-134 verbatim
    w = rbm.weights + torch.zeros((z_vis.size(0),) + rbm.weights.size(), device=DEV)
    vis = z_vis.unsqueeze(2).permute(0, 2, 1); hid = z_hid.unsqueeze(2)
    ref = torch.mean(-torch.matmul(vis, torch.matmul(w, hid)).reshape(-1) - torch.matmul(z_vis, rbm.visible_bias)
                     - torch.matmul(z_hid, rbm.hidden_bias), 0)
    gr = torch.autograd.grad(ref, [z_vis, z_hid, rbm._weights, rbm._visible_bias, rbm._hidden_bias])
    torch.testing.assert_close(e, ref, rtol=1e-4, atol=1e-4)
    for mine, r in zip(g, gr):
        torch.testing.assert_close(mine, r, rtol=1e-3, atol=1e-5)


def test_sample_many_chains_in_one_launch():
    """generation of num_samples chains at once instead of a loop of batch-size calls."""
    W, a, c = rand_params(64, 64, 0.3, 22)
    rbm = make_rbm(W, a, c)
    s = divae_b200.PCD(batch_size=128, RBM=rbm, n_gibbs_sampling_steps=10, seed=31)
    v, h = s.sample(1024)
    assert v.shape == (1024, 64) and h.shape == (1024, 64)
    ov, oh, clean = orc.block_gibbs_philox(W, a, c, 1024, 10, 31, chain_offset=128)
    assert (v.cpu().numpy()[clean] == ov[clean]).all() and (h.cpu().numpy()[clean] == oh[clean]).all()
    v2, h2 = s.block_gibbs_sampling()                       # the sampler's own batch is untouched
    assert v2.shape == (128, 64) and s.get_batch_size() == 128
    ov2, oh2, clean2 = orc.block_gibbs_philox(W, a, c, 128, 10, 31, step_offset=20)
    assert (v2.cpu().numpy()[clean2] == ov2[clean2]).all()


def test_statistics_and_decoder_entries_are_cuda_graph_capturable():
    """rbm_b200_block_gibbs_stats on the tensor-core path (statistics GEMM, cluster-launched energy reduction)
    and rbm_b200_decoder_forward: captured on a side stream, replayed, results equal the eager call."""
    from divae_b200.decoder import DecoderRunner
    side = torch.cuda.Stream()
    rbm = divae_b200.RBM(320, 256).to(DEV)
    s = divae_b200.PCD(batch_size=700, RBM=rbm, n_gibbs_sampling_steps=2, seed=3, variant="umma", persistent=False)

    class V3(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self._activation_fct = torch.nn.ReLU()
            self._layers = torch.nn.ModuleList([torch.nn.Linear(33, 40), torch.nn.Linear(40, 56)])
            self._layers2 = torch.nn.ModuleList([torch.nn.Linear(56, 72), torch.nn.Linear(72, 100)])
            self._layers3 = torch.nn.ModuleList([torch.nn.Linear(56, 72), torch.nn.Linear(72, 100)])

    torch.manual_seed(1)
    r = DecoderRunner(V3().to(DEV))
    x = torch.randn(300, 33, device=DEV)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        e0, (S0, sv0, sh0) = s.negative_phase()          # warm-up: attributes, workspaces, staged parameters
        d0 = r.shower(x, seed=4)
        side.synchronize()
        half_steps = s._half_steps
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            e1, (S1, sv1, sh1) = s.negative_phase()
            d1 = r.shower(x, seed=4)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(d0, d1)
    # the captured sampler call used the next half-step counters: compare with an eager sampler at that offset
    s2 = divae_b200.PCD(batch_size=700, RBM=rbm, n_gibbs_sampling_steps=2, seed=3, variant="umma")
    s2._half_steps = half_steps
    e2, (S2, sv2, sh2) = s2.negative_phase()
    assert torch.equal(S1, S2) and torch.equal(sv1, sv2) and torch.equal(sh1, sh2)
    assert abs(e1.item() - e2.item()) <= 1e-5 * max(1.0, abs(e2.item()))


def test_kernel_choice_by_shape():
    """The shape-based choice inside the library, as bench.py reports it."""
    lib = _lib.load()
    name = lambda nV, nH, B, variant: lib.rbm_b200_kernel_choice(nV, nH, B, _lib.VARIANTS[variant]).decode()
    assert name(64, 64, 100, "auto") == "smem_gibbs_kernel"                      # BASELINE configs[1]
    assert name(256, 256, 1024, "auto") == "smem_gibbs_kernel"                   # configs[2]
    assert name(1024, 1024, 65536, "auto") == "umma_half_step_kernel"            # configs[3] on one GPU
    assert name(1024, 1024, 8192, "auto") == "umma_chain_kernel"                 # configs[3] sharded over 8 GPUs
    assert name(512, 512, 2048, "umma") == "umma_resident_kernel"                # configs[4] shard at 8 GPUs
    assert name(512, 512, 16384, "umma") == "umma_chain_kernel"                  # configs[4] on one GPU
    assert name(512, 512, 2048, "umma_split") == "umma_chain_kernel"             # split weights: chain kernel only
    assert name(1024, 1024, 64, "auto") in ("umma_chain_kernel", "umma_half_step_kernel")   # never the fp32 GENERAL path
