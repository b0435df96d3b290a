"""Shower decoding throughput: decoder MLP + shower tail on generated prior samples.

Shape of configs/model/gumboltcaloV.yaml: latent 2*64+1 -> 200 -> 300 -> [400 -> 504] x 2 heads.
Reports, for B samples on one GPU: this library's fused path (rbm_b200_decoder_forward, samples only), the same
network as torch.nn modules in fp32 on the GPU (ATen/cuBLAS, TF32 off: the reference's own code path on a GPU),
and the whole generate_samples() (sampler k = 40 + decoder)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import divae_b200  # noqa: E402
from divae_b200.decoder import DecoderRunner, generate_samples  # noqa: E402


class V3(torch.nn.Module):
    def __init__(self, nodes):
        super().__init__()
        self._activation_fct = torch.nn.ReLU()
        self._output_activation_fct = torch.nn.Identity()
        n = len(nodes)
        self._layers = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in nodes[:n - 2]])
        self._layers2 = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in nodes[n - 2:]])
        self._layers3 = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in nodes[n - 2:]])

    def forward(self, x):   # BasicDecoderV3.forward (basicCoders.py:68-84)
        for layer in self._layers:
            x = torch.relu(layer(x))
        x1, x2 = x, x
        n = len(self._layers2)
        for i, (l2, l3) in enumerate(zip(self._layers2, self._layers3)):
            x1, x2 = l2(x1), l3(x2)
            if i != n - 1:
                x1, x2 = torch.relu(x1), torch.relu(x2)
        return x1, x2


def timeit(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    nodes = [(129, 200), (200, 300), (300, 400), (400, 504)]
    m = V3(nodes).cuda()
    r = DecoderRunner(m)
    x = torch.cat([(torch.rand(B, 128, device="cuda") < 0.5).float(), torch.rand(B, 1, device="cuda") * 100], 1)
    flops = 2.0 * B * (129 * 200 + 200 * 300 + 2 * (300 * 400 + 400 * 504))

    def torch_path():
        with torch.no_grad():
            hits, act = m(x)
            rho = torch.rand_like(hits)
            return torch.relu(act) * torch.heaviside(hits + torch.log(rho) - torch.log(1 - rho), torch.zeros(1, device="cuda"))

    out = {"samples": B, "nodes": nodes, "flops_fp32_equiv": flops}
    out["b200_shower_ms"] = timeit(lambda: r.shower(x, seed=1))
    out["b200_forward_ms"] = timeit(lambda: r(x))
    out["torch_fp32_gpu_ms"] = timeit(torch_path)
    out["b200_showers_per_s"] = B / (out["b200_shower_ms"] * 1e-3)
    out["b200_tflops_fp32_equiv"] = flops / (out["b200_shower_ms"] * 1e-3) / 1e12

    class Model:
        pass
    model = Model()
    model.sampler = divae_b200.PCD(batch_size=128, RBM=divae_b200.RBM(64, 64).cuda(), n_gibbs_sampling_steps=40, seed=3)
    model.decoder = m
    out["generate_samples_ms"] = timeit(lambda: generate_samples(model, num_samples=B, true_energy=50.0, seed=8))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
