"""Validation-style forward pass of the calorimeter model (gumboltcaloV.yaml shapes) on the B200 path:
HierarchicalEncoder in evaluation mode (EncoderRunner) -> decoder + shower tail (DecoderRunner), against the same
networks as ATen fp32 modules on the same GPU (the reference's code path, engineCaloV3 validation)."""
import json
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchmarks.decoder_bench import V3, timeit  # noqa: E402
from divae_b200.decoder import DecoderRunner, EncoderRunner  # noqa: E402


class Encoder(torch.nn.Module):
    def __init__(self, n_in, n_latent, hidden0, hidden, beta):
        super().__init__()
        self._config = SimpleNamespace(model=SimpleNamespace(beta_smoothing_fct=beta))
        nets = []
        for lvl in range(2):
            sizes = ([n_in] + hidden0 + [n_latent]) if lvl == 0 else ([n_in + lvl * n_latent] + hidden + [n_latent])
            mods = []
            for i in range(len(sizes) - 1):
                mods.append(torch.nn.Linear(sizes[i], sizes[i + 1]))
                mods.append(torch.nn.Identity() if i == len(sizes) - 2 else torch.nn.ReLU())
            nets.append(torch.nn.Sequential(*mods))
        self._networks = torch.nn.ModuleList(nets)


def main():
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    enc = Encoder(504, 64, [400, 300, 200], [64, 64], 7.0).cuda()
    dec = V3([(129, 200), (200, 300), (300, 400), (400, 504)]).cuda()
    er, dr = EncoderRunner(enc), DecoderRunner(dec)
    x = torch.relu(torch.randn(B, 504, device="cuda"))
    e = torch.rand(B, 1, device="cuda") * 100

    def b200():
        _, _, z = er(x, is_training=False, seed=1)
        return dr.shower(torch.cat(z + [e], 1), seed=2)

    def aten():
        with torch.no_grad():
            zs = []
            for net in enc._networks:
                l = torch.clamp(net(torch.cat([x] + zs, 1)), -88., 88.)
                rho = torch.rand_like(l)
                zs.append(torch.heaviside(l + torch.log(rho) - torch.log(1 - rho), torch.zeros(1, device="cuda")))
            hits, act = dec(torch.cat(zs + [e], 1))
            rho = torch.rand_like(hits)
            return torch.relu(act) * torch.heaviside(hits + torch.log(rho) - torch.log(1 - rho), torch.zeros(1, device="cuda"))

    out = {"samples": B, "b200_ms": timeit(b200), "aten_fp32_ms": timeit(aten)}
    out["b200_samples_per_s"] = B / (out["b200_ms"] * 1e-3)
    out["speedup"] = out["aten_fp32_ms"] / out["b200_ms"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
