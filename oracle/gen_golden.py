"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE (build container only).

TEST INFRASTRUCTURE.  Run as `python -m oracle.gen_golden` from the repo root in
a container that has /root/reference.  The reference has no uniform-injection
hook, so `torch.rand` is interposed while its code runs: the
sampler then consumes the exact uniforms of the device draw contract
(oracle/philox.py) as float64 tensors — `probs >= u` promotes to float64, so
the comparison is exact.  Outputs are what the reference's own code returned.

Fixtures (all small):
  pcd_*.npz        PCD.hidden_samples / visible_samples / block_gibbs_sampling  (pcd.py:23-69)
  energy.npz       RBM.energy, log_prob, cross_entropy                            (rbm.py:56-80)
  meanfield.npz    GibbsSampler.gibbs_sampling / get_samples                      (gibbsSampler.py:31-63)
  cd1.npz          sandbox Contrastive_Divergence.contrastive_divergence          (rbm_standalone.py:75-121)
  negphase.npz     GumBolt.energy_exp + its autograd w.r.t. (W,a,c)               (gumbolt.py:109-134)
  port_check.npz   torch port == reference under the same torch seed (bit-exact)
  decoder_v3.npz   BasicDecoderV3.forward + the shower tail of generate_samples with GumbelMod
                   (basicCoders.py:63-84, gumboltCaloV5.py:88-93, gumbelmod.py:18-24)
  decoder_basic.npz  BasicDecoder.forward (basicCoders.py:28-42)
  encoder.npz      HierarchicalEncoder.forward with GumbelMod, training and evaluation mode
                   (hierarchicalEncoder.py:139-183, gumbelmod.py:14-24)
"""
import os
import sys

import numpy as np
import torch

from . import _ref_shim, decoder_oracle as dec, rbm_oracle as orc, ref_torch_port as port

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class RandInterposer:
    """Context manager: torch.rand(...) hands out queued tensors instead of drawing."""

    def __init__(self, supplier):
        self.supplier = supplier
        self.calls = 0

    def __enter__(self):
        self._orig = torch.rand

        def fake(*size, **kw):
            if len(size) == 1 and not isinstance(size[0], int):
                size = tuple(size[0])
            t = self.supplier(self.calls, tuple(size))
            self.calls += 1
            return t

        torch.rand = fake
        return self

    def __exit__(self, *exc):
        torch.rand = self._orig


def packbits(x):
    return np.packbits(np.asarray(x) > 0.5, axis=1)


def make_params(nV, nH, scale, seed, bias_sigma=0.2):
    rng = np.random.default_rng(seed)
    W = orc.round_bf16((rng.standard_normal((nV, nH)) * scale).astype(np.float32))
    a = (0.5 + bias_sigma * rng.standard_normal(nV)).astype(np.float32)
    c = (bias_sigma * rng.standard_normal(nH)).astype(np.float32)
    return W, a, c


def ref_rbm(W, a, c):
    from models.rbm.rbm import RBM
    r = RBM(n_visible=W.shape[0], n_hidden=W.shape[1])
    with torch.no_grad():
        r._weights.copy_(torch.from_numpy(W))
        r._visible_bias.copy_(torch.from_numpy(a))
        r._hidden_bias.copy_(torch.from_numpy(c))
    return r


def gen_pcd(name, nV, nH, B, k, scale, seed):
    from models.samplers.pcd import PCD
    W, a, c = make_params(nV, nH, scale, seed)
    rbm = ref_rbm(W, a, c)
    v0 = orc.init_state_philox(B, nV, seed)
    draw = orc.philox_uniforms(seed, B)

    # constructor draws two rand(B,nV) (pcd.py:16-17): feed anything, then overwrite _MCState
    with RandInterposer(lambda i, s: torch.zeros(s)):
        sampler = PCD(batch_size=B, RBM=rbm, n_gibbs_sampling_steps=k)
    sampler._MCState = torch.from_numpy(v0)

    def supplier(i, shape):
        if i < 2 * k:  # step s: rand([B,nH]) then rand([B,nV])
            return torch.from_numpy(draw(i, shape))
        return torch.zeros(shape)  # the two re-init draws (pcd.py:66-67)

    with RandInterposer(supplier) as ri, torch.no_grad():
        v, h = sampler.block_gibbs_sampling()
    assert ri.calls == 2 * k + 2, ri.calls
    # single half-steps
    with RandInterposer(lambda i, s: torch.from_numpy(draw(0, s))), torch.no_grad():
        h1 = sampler.hidden_samples(torch.from_numpy(v0))
    with RandInterposer(lambda i, s: torch.from_numpy(draw(1, s))), torch.no_grad():
        v1 = sampler.visible_samples(h1)
    # oracle agrees with the reference on every chain that never met a near-threshold draw
    ov, oh, clean = orc.block_gibbs_philox(W, a, c, B, k, seed)
    vn, hn = v.numpy(), h.numpy()
    assert (ov[clean] == vn[clean]).all() and (oh[clean] == hn[clean]).all(), name
    np.savez_compressed(os.path.join(OUT, name + ".npz"), W=W, a=a, c=c, B=B, k=k, seed=seed,
                        v0=packbits(v0), v=packbits(vn), h=packbits(hn),
                        h1=packbits(h1.numpy()), v1=packbits(v1.numpy()))
    print(name, "clean chains %d/%d" % (clean.sum(), B), "mean v %.3f h %.3f" % (vn.mean(), hn.mean()))


def gen_energy():
    W, a, c = make_params(48, 48, 0.3, 11)
    rbm = ref_rbm(W, a, c)
    rng = np.random.default_rng(5)
    zb = (rng.random((40, 96)) < 0.5).astype(np.float32)
    zr = rng.random((40, 96)).astype(np.float32)
    out = {}
    with torch.no_grad():
        for nm, z in (("bin", zb), ("real", zr)):
            parts = [torch.from_numpy(z[:, :30].copy()), torch.from_numpy(z[:, 30:].copy())]  # ragged list, cat'ed
            out["E_" + nm] = rbm.energy(parts).numpy()
            out["logp_" + nm] = rbm.log_prob(parts).numpy()
            out["xent_" + nm] = rbm.cross_entropy(parts).numpy()
    for nm, z in (("bin", zb), ("real", zr)):
        e = orc.energy([z[:, :30], z[:, 30:]], W, a, c)
        assert np.allclose(e, out["E_" + nm], rtol=1e-5, atol=1e-5)
    np.savez_compressed(os.path.join(OUT, "energy.npz"), W=W, a=a, c=c, zb=zb, zr=zr,
                        logZ=np.float32(rbm.get_logZ_value()), **out)
    print("energy ok")


def gen_meanfield():
    from models.samplers.gibbsSampler import GibbsSampler
    W, a, c = make_params(200, 200, 0.05, 21)  # divae.yaml shape (4 x 100 latent nodes)
    rbm = ref_rbm(W, a, c)
    gs = GibbsSampler(RBM=rbm, n_gibbs_sampling_steps=40)
    rng = np.random.default_rng(3)
    x = rng.random((16, 200)).astype(np.float32)
    with torch.no_grad():
        left, right = gs.gibbs_sampling(torch.from_numpy(x), 7)
    # get_samples with the default random start -166*rand+88 (gibbsSampler.py:57-61)
    starts = rng.random((2, 100)).astype(np.float32)
    with RandInterposer(lambda i, s: torch.from_numpy(starts[i])), torch.no_grad():
        gl, gr = gs.get_samples(approx_post_samples=[], n_latent_nodes=100, n_latent_hierarchy_lvls=4,
                                n_gibbs_sampling_steps=5)
    ol, orr = orc.meanfield_gibbs(x, W, a, c, 7)
    assert np.allclose(ol, left.numpy(), atol=2e-6) and np.allclose(orr, right.numpy(), atol=2e-6)
    np.savez_compressed(os.path.join(OUT, "meanfield.npz"), W=W, a=a, c=c, x=x, k=7,
                        left=left.numpy(), right=right.numpy(), starts=starts, gk=5,
                        gleft=gl.numpy(), gright=gr.numpy())
    print("meanfield ok")


def gen_cd1():
    sys.path.insert(0, os.path.join(_ref_shim.REF_ROOT, "sandbox"))
    import rbm_standalone
    nV, nH, B, k = 784, 128, 128, 1
    rng = np.random.default_rng(7)
    W, a, c = make_params(nV, nH, 0.01, 7, bias_sigma=0.0)
    v0 = (rng.random((B, nV)) > 0.5).astype(np.float32)
    U = rng.random((k + 1, nH)).astype(np.float32)
    cd = rbm_standalone.Contrastive_Divergence(n_visible=nV, n_hidden=nH, learning_rate=1e-3,
                                               momentum_coefficient=0.5, n_gibbs_sampling_steps=k,
                                               weight_cost=1e-4)
    cd._weights = torch.from_numpy(W.copy()); cd._visible_bias = torch.from_numpy(a.copy())
    cd._hidden_bias = torch.from_numpy(c.copy())
    with RandInterposer(lambda i, s: torch.from_numpy(U[i])):
        loss = cd.contrastive_divergence(torch.from_numpy(v0))
    st = orc.cd_k(v0, W, a, c, k, lambda i, s: U[i][None, :])
    state = dict(W=W.copy(), a=a.copy(), c=c.copy(), dW=np.zeros_like(W), da=np.zeros_like(a), dc=np.zeros_like(c))
    state = orc.cd_update(state, st)
    assert np.allclose(state["W"], cd._weights.numpy(), atol=1e-6)
    assert np.allclose(state["a"], cd._visible_bias.numpy(), atol=1e-5)
    assert np.allclose(state["c"], cd._hidden_bias.numpy(), atol=1e-5)
    assert np.allclose(st["recon_sse"], loss.item(), rtol=1e-5)
    np.savez_compressed(os.path.join(OUT, "cd1.npz"), W=W, a=a, c=c, v0=packbits(v0), U=U, k=k,
                        W_new=cd._weights.numpy().astype(np.float16), a_new=cd._visible_bias.numpy(),
                        c_new=cd._hidden_bias.numpy(), dW=cd.weights_update.numpy().astype(np.float32)[::8, ::8],
                        loss=np.float32(loss.item()))
    print("cd1 ok, loss", loss.item())


def gen_negphase():
    from types import SimpleNamespace
    from models.autoencoders.gumbolt import GumBolt
    W, a, c = make_params(64, 64, 0.3, 31)
    rbm = ref_rbm(W, a, c)
    rng = np.random.default_rng(9)
    v = (rng.random((100, 64)) < 0.6).astype(np.float32)
    h = (rng.random((100, 64)) < 0.4).astype(np.float32)
    fake_self = SimpleNamespace(prior=rbm)
    e = GumBolt.energy_exp(fake_self, torch.from_numpy(v), torch.from_numpy(h))
    gW, ga, gc = torch.autograd.grad(-e, [rbm._weights, rbm._visible_bias, rbm._hidden_bias])
    S, sv, sh = orc.negative_phase_stats(v, h)
    assert np.allclose(S, gW.numpy(), atol=1e-6) and np.allclose(sv, ga.numpy(), atol=1e-6)
    assert np.allclose(sh, gc.numpy(), atol=1e-6)
    assert np.allclose(orc.energy_exp(v, h, W, a, c), e.item(), rtol=1e-5)
    np.savez_compressed(os.path.join(OUT, "negphase.npz"), W=W, a=a, c=c, v=packbits(v), h=packbits(h),
                        e=np.float32(e.item()), gW=gW.numpy(), ga=ga.numpy(), gc=gc.numpy())
    print("negphase ok", e.item())


def gen_port_check():
    """torch port (oracle/ref_torch_port.py) is bit-identical to the reference under one torch seed."""
    from models.samplers.pcd import PCD
    W, a, c = make_params(64, 64, 0.3, 41)
    rbm = ref_rbm(W, a, c)
    torch.manual_seed(32)
    s = PCD(batch_size=100, RBM=rbm, n_gibbs_sampling_steps=10)
    with torch.no_grad():
        v, h = s.block_gibbs_sampling()
    torch.manual_seed(32)
    st = port.pcd_init_state(100, 64)
    with torch.no_grad():
        pv, ph, st2 = port.block_gibbs_sampling(st, torch.from_numpy(W), torch.from_numpy(a), torch.from_numpy(c), 10)
    assert torch.equal(v, pv) and torch.equal(h, ph) and torch.equal(st2, s._MCState)
    np.savez_compressed(os.path.join(OUT, "port_check.npz"), W=W, a=a, c=c, torch_seed=32, B=100, k=10,
                        v=packbits(v.numpy()), h=packbits(h.numpy()), torch_version=torch.__version__)
    print("port bit-identical to reference PCD")


def _layer_arrays(mods):
    return [(m.weight.detach().numpy().copy(), m.bias.detach().numpy().copy()) for m in mods]


def gen_decoder():
    """The reference's own decoder classes and GumbelMod on a small node sequence (the structure - trunk, two
    heads, where ReLU and Identity apply, the heaviside gate - is what the oracle has to match)."""
    from models.networks.basicCoders import BasicDecoder, BasicDecoderV3
    from utils.dists.gumbelmod import GumbelMod
    torch.manual_seed(32)
    rng = np.random.default_rng(77)
    B, seed = 24, 0x5EED + 9
    # --- two-headed calorimeter decoder: latent 2n + 1 -> trunk -> (hits, activations)
    nodes = [(33, 40), (40, 56), (56, 72), (72, 100)]
    d3 = BasicDecoderV3(node_sequence=nodes, activation_fct=torch.nn.ReLU(), cfg=None)
    states = (rng.random((B, 32)) < 0.5).astype(np.float32)
    true_e = (rng.random((B, 1)) * 100).astype(np.float32)
    x = np.concatenate([states, true_e], axis=1)
    with torch.no_grad():
        hits, act = d3(torch.from_numpy(x))
    u = dec.hit_uniforms(seed, B, 100)
    rho = (1.0 - u).astype(np.float32)
    with RandInterposer(lambda i, s_: torch.from_numpy(rho)) as ri, torch.no_grad():
        gate = GumbelMod()(hits, torch.tensor(7.0), False)
        sample = torch.relu(act) * gate                              # gumboltCaloV5.py:93
    assert ri.calls == 1
    trunk, hh, ha = _layer_arrays(d3._layers), _layer_arrays(d3._layers2), _layer_arrays(d3._layers3)
    oh, oa = dec.decoder_v3(trunk, hh, ha, x)
    np.testing.assert_allclose(oh, hits.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(oa, act.numpy(), rtol=1e-5, atol=1e-5)
    assert np.array_equal(dec.shower_tail(hits.numpy(), act.numpy(), rho), sample.numpy())
    so, clean = dec.shower_philox(hits.numpy(), act.numpy(), seed)
    assert clean.mean() > 0.99 and np.array_equal(so[clean], sample.numpy()[clean])
    flat = {}
    for name, ch in (("trunk", trunk), ("hits", hh), ("act", ha)):
        for i, (W, b) in enumerate(ch):
            flat["%s_W%d" % (name, i)] = W; flat["%s_b%d" % (name, i)] = b
    np.savez_compressed(os.path.join(OUT, "decoder_v3.npz"), x=x, out_hits=hits.numpy(), out_act=act.numpy(), rho=rho,
                        sample=sample.numpy(), seed=seed, n_trunk=len(trunk), n_head=len(hh), **flat)
    print("decoder_v3: hit rate %.3f, clean %.4f" % (gate.numpy().mean(), clean.mean()))
    # --- single-chain decoder (GumBolt / DiVAE MNIST-shape models)
    nodes = [(16, 24), (24, 40), (40, 50)]
    d1 = BasicDecoder(node_sequence=nodes, activation_fct=torch.nn.ReLU(), output_activation_fct=torch.nn.Identity(), cfg=None)
    x1 = rng.standard_normal((B, 16)).astype(np.float32)
    with torch.no_grad():
        y1 = d1(torch.from_numpy(x1)).numpy()
    layers = _layer_arrays(d1._layers)
    np.testing.assert_allclose(dec.decoder_basic(layers, x1), y1, rtol=1e-5, atol=1e-5)
    flat = {}
    for i, (W, b) in enumerate(layers):
        flat["W%d" % i] = W; flat["b%d" % i] = b
    np.savez_compressed(os.path.join(OUT, "decoder_basic.npz"), x=x1, y=y1, n_layers=len(layers), **flat)
    print("decoder_basic ok")


def gen_encoder():
    """The reference's HierarchicalEncoder (Gumbel smoother) fed the uniforms of the device contract."""
    from types import SimpleNamespace
    from models.networks.hierarchicalEncoder import HierarchicalEncoder
    torch.manual_seed(33)
    rng = np.random.default_rng(5)
    B, n_in, n_lat, seed, beta = 20, 50, 16, 0x5EED + 11, 7.0
    cfg = SimpleNamespace(model=SimpleNamespace(encoder_hidden_nodes=[40, 30], beta_smoothing_fct=beta))
    enc = HierarchicalEncoder(activation_fct=torch.nn.ReLU(), input_dimension=n_in, n_latent_hierarchy_lvls=2,
                              n_latent_nodes=n_lat, n_encoder_layer_nodes=24, n_encoder_layers=2, skip_latent_layer=False,
                              smoother="Gumbel", cfg=cfg)
    x = rng.standard_normal((B, n_in)).astype(np.float32)
    levels = [[(m.weight.detach().numpy().copy(), m.bias.detach().numpy().copy()) for m in net if isinstance(m, torch.nn.Linear)]
              for net in enc._networks]
    out = dict(x=x, beta=beta, seed=seed, n_levels=len(levels), n_layers=len(levels[0]))
    for li, ch in enumerate(levels):
        for i, (W, b) in enumerate(ch):
            out["L%d_W%d" % (li, i)] = W; out["L%d_b%d" % (li, i)] = b
    for mode, is_training in (("train", True), ("eval", False)):
        # training: the device's fp32 rho; evaluation: rho = 1 - u with the exact u (float64 promotes the compare)
        def rho(lvl, shape):
            u32 = dec.smooth_u32(seed, shape[0], shape[1], 0, 0, lvl * n_lat)
            if is_training:
                return dec.smooth_rho(u32)
            return (1.0 - u32.astype(np.float64) / 4294967296.0).astype(np.float32)
        calls = []
        with RandInterposer(lambda i, s_: (calls.append(s_), torch.from_numpy(rho(i, s_)))[1]), torch.no_grad():
            b_, logits, samples = enc(torch.from_numpy(x), is_training=is_training)
        assert len(calls) == 2 and float(b_) == beta
        ol, os_ = dec.encoder_forward(levels, x, beta, is_training, rho)
        for lvl in range(2):
            np.testing.assert_allclose(ol[lvl], logits[lvl].numpy(), rtol=1e-5, atol=1e-5)
            if is_training:
                np.testing.assert_allclose(os_[lvl], samples[lvl].numpy(), rtol=0, atol=2e-5)
            out["%s_logits%d" % (mode, lvl)] = logits[lvl].numpy(); out["%s_samples%d" % (mode, lvl)] = samples[lvl].numpy()
        pl, ps, clean = dec.encoder_forward_philox(levels, x, beta, is_training, seed)
        if not is_training:
            assert clean.mean() > 0.9
            for lvl in range(2):
                assert np.array_equal(ps[lvl][clean], samples[lvl].numpy()[clean]), lvl
        print("encoder %s: mean sample %.3f" % (mode, np.mean([s_.numpy().mean() for s_ in samples])), "clean %.3f" % clean.mean())
    np.savez_compressed(os.path.join(OUT, "encoder.npz"), **out)


def main():
    _ref_shim.install()
    os.makedirs(OUT, exist_ok=True)
    gen_pcd("pcd_64x64", 64, 64, 100, 10, 0.3, 0x5EED)          # BASELINE config 2
    gen_pcd("pcd_256x256", 256, 256, 64, 20, 0.1, 0x5EED + 1)   # BASELINE config 3 prior, fewer chains
    gen_pcd("pcd_784x128", 784, 128, 32, 3, 0.05, 0x5EED + 2)   # config 1 prior shape (ragged nV != nH)
    gen_pcd("pcd_200x200", 200, 200, 24, 4, 0.1, 0x5EED + 3)    # divae.yaml prior (not a multiple of 16)
    gen_pcd("pcd_refinit_64x64", 64, 64, 100, 10, 0.01, 0x5EED + 4)  # reference init scale (rbm.py:28)
    gen_energy()
    gen_meanfield()
    gen_cd1()
    gen_negphase()
    gen_port_check()
    gen_decoder()
    gen_encoder()


if __name__ == "__main__":
    main()
