"""DiVAE (GumBolt) train samples/s with the B200 sampler vs the reference's ATen sampler, same GPU.

BASELINE.json's second metric ("DiVAE train samples/s"; SURVEY.md §8d "secondary").  The reference's training
loop cannot travel to the GPU box, so this is a compact GumBolt-shaped model written in plain torch (hierarchical
encoder with Gumbel smoothing, MLP decoder, KL = entropy + positive energy + negative energy, Adam), i.e. the
same op mix as models/autoencoders/gumbolt.py:36-134 around the prior.  Only the prior side differs between arms:

  aten  the reference's own formulation on the GPU: a Python loop of matmul / sigmoid / rand / compare per
        half-step (models/samplers/pcd.py:23-69) and energy_exp with W broadcast to [B, nV, nH] (gumbolt.py:109-134)
  b200  PCD.negative_phase() (one sampling launch + fused statistics) and divae_b200.energy_exp

Shapes: BASELINE configs[1] (RBM 64x64, 100 chains x 10 steps, batch 100), the reference default PCD(128, k=40),
and configs[2] (RBM 256x256, 1024 chains x 20 steps, 504 voxels).
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import divae_b200  # noqa: E402


class GumBoltLike(torch.nn.Module):
    def __init__(self, n_in, n_latent, enc_hidden, dec_hidden, beta=5.0):
        super().__init__()
        self.beta = beta
        self.n_latent = n_latent

        def mlp(sizes):
            layers = []
            for i in range(len(sizes) - 1):
                layers.append(torch.nn.Linear(sizes[i], sizes[i + 1]))
                if i != len(sizes) - 2:
                    layers.append(torch.nn.ReLU())
            return torch.nn.Sequential(*layers)

        self.enc0 = mlp([n_in] + enc_hidden + [n_latent])
        self.enc1 = mlp([n_in + n_latent] + enc_hidden + [n_latent])
        self.dec = mlp([2 * n_latent] + dec_hidden + [n_in])
        self.prior = divae_b200.RBM(n_latent, n_latent)

    def posterior(self, x):
        logits, zetas = [], []
        inp = x
        for net in (self.enc0, self.enc1):
            l = torch.clamp(net(inp), -88., 88.)
            rho = torch.rand_like(l)
            z = torch.sigmoid((l + torch.log(rho) - torch.log(1 - rho)) * self.beta)   # Gumbel smoothing
            logits.append(l); zetas.append(z)
            inp = torch.cat([x, z], 1)
        return torch.cat(logits, 1), torch.cat(zetas, 1)


def energy_exp_broadcast(prior, vis, hid):
    """The reference's expression (gumbolt.py:109-134): W broadcast to [B, nV, nH]."""
    w = prior.weights + torch.zeros((vis.size(0),) + prior.weights.size(), device=vis.device)
    e = (-torch.matmul(vis.unsqueeze(1), torch.matmul(w, hid.unsqueeze(2))).reshape(-1)
         - torch.matmul(vis, prior.visible_bias) - torch.matmul(hid, prior.hidden_bias))
    return e.mean(0)


def aten_block_gibbs(prior, B, k):
    """pcd.py:51-69 on the GPU: fresh Bernoulli(1/2) chains, k x (h | v, v | h) as separate ATen ops."""
    W, a, c = prior.weights, prior.visible_bias, prior.hidden_bias
    v = (torch.rand(B, W.shape[0], device=W.device) >= torch.rand(B, W.shape[0], device=W.device)).float()
    h = None
    for _ in range(k):
        ph = torch.sigmoid(torch.matmul(v, W) + c)
        h = (ph >= torch.rand(ph.size(), device=W.device)).float()
        pv = torch.sigmoid(torch.matmul(h, W.t()) + a)
        v = (pv >= torch.rand(pv.size(), device=W.device)).float()
    return v, h


def train_step(model, opt, sampler, x, arm, chains, k):
    logits, zetas = model.posterior(x)
    recon = model.dec(zetas)
    ae = torch.nn.functional.binary_cross_entropy_with_logits(recon, x, reduction="none").sum(1).mean()
    entropy = -torch.nn.functional.binary_cross_entropy_with_logits(logits, zetas, reduction="none").sum(1).mean()
    n = model.n_latent
    zv, zh = zetas[:, :n], zetas[:, n:]
    if arm == "aten":
        pos = energy_exp_broadcast(model.prior, zv, zh)
        v, h = aten_block_gibbs(model.prior, chains, k)        # grad-enabled like engineDiVAEpp.py:40
        neg = -energy_exp_broadcast(model.prior, v.detach(), h.detach())
    else:
        pos = divae_b200.energy_exp(model.prior, zv, zh)
        e_neg, _ = sampler.negative_phase()
        neg = -e_neg
    loss = ae + entropy + pos + neg
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss


def run(name, n_in, n_latent, enc_hidden, dec_hidden, batch, chains, k, iters=50):
    out = {"case": name, "batch": batch, "chains": chains, "gibbs_steps": k, "rbm": "%dx%d" % (n_latent, n_latent)}
    for arm in ("aten", "b200"):
        torch.manual_seed(0)
        model = GumBoltLike(n_in, n_latent, enc_hidden, dec_hidden).cuda()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        sampler = divae_b200.PCD(batch_size=chains, RBM=model.prior, n_gibbs_sampling_steps=k, seed=1)
        x = (torch.rand(batch, n_in, device="cuda") > 0.5).float()
        for _ in range(5):
            train_step(model, opt, sampler, x, arm, chains, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            loss = train_step(model, opt, sampler, x, arm, chains, k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out[arm + "_ms_per_step"] = ms
        out[arm + "_train_samples_per_s"] = batch / (ms * 1e-3)
        out[arm + "_loss"] = float(loss.detach())
    out["speedup"] = out["aten_ms_per_step"] / out["b200_ms_per_step"]
    print(json.dumps(out), flush=True)


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    # BASELINE configs[1]: MNIST shape, RBM 2 x 64 latent units, 100 chains x 10 steps per iteration
    run("config2_mnist_64x64_pcd100x10", 784, 64, [64, 64, 64], [64, 64, 64], 100, 100, 10)
    # the sampler the reference's DiVAE++/GumBolt models actually build: PCD(128, k = 40) (dvaepp.py:57)
    run("reference_default_64x64_pcd128x40", 784, 64, [64, 64, 64], [64, 64, 64], 128, 128, 40)
    # BASELINE configs[2]: calorimeter shape (504 voxels), RBM 2 x 256, 1024 chains x 20 steps
    run("config3_calo_256x256_pcd1024x20", 504, 256, [400, 300, 200], [200, 300, 400], 128, 1024, 20, iters=20)


if __name__ == "__main__":
    main()
