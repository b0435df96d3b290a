"""Training step of the posterior side (hierarchical encoder + Gumbel smoother, forward + backward) on the tcgen05 GEMMs
against the same torch modules (ATen / cuBLAS fp32), calorimeter shape 504 -> 400 -> 300 -> 200 -> 64, 568 -> 64 -> 64 -> 64.
Also the raw GEMM primitives against torch.matmul.  Usage: python benchmarks/backward_bench.py [batch ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from divae_b200 import _ops  # noqa: E402
from divae_b200.decoder import EncoderRunner  # noqa: E402
from test_gpu_encoder import Encoder  # noqa: E402

DEV = "cuda:0"


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for B in [int(a) for a in sys.argv[1:]] or [1024, 65536]:
    torch.manual_seed(0)
    enc = Encoder(504, 64, 2, [400, 300, 200], [64, 64], 7.0).to(DEV)
    runner = EncoderRunner(enc)
    x = torch.rand(B, 504, device=DEV)

    def ours():
        _, lg, z = runner(x, is_training=True, seed=1)
        (z[0].sum() + z[1].sum() + lg[1].sum()).backward()

    def aten():
        post, lgs = [], []
        for net in enc._networks:
            l = torch.clamp(net(torch.cat([x] + post, 1)), -88., 88.)
            rho = torch.rand_like(l)
            post.append(torch.sigmoid((l + torch.log(rho) - torch.log(1 - rho)) * 7.0)); lgs.append(l)
        (post[0].sum() + post[1].sum() + lgs[1].sum()).backward()

    a = torch.randn(B, 1024, device=DEV); w = torch.randn(1024, 1024, device=DEV)
    flops = 2.0 * B * 1024 * 1024
    line = {"batch": B, "encoder_fwd_bwd_ms_b200": timed(ours), "encoder_fwd_bwd_ms_aten": timed(aten),
            "gemm_nt_1024_ms": timed(lambda: _ops.gemm(a, w, True)), "matmul_nt_1024_ms": timed(lambda: a @ w.t()),
            "gemm_tn_1024_ms": timed(lambda: _ops.gemm_tn(a, a)), "matmul_tn_1024_ms": timed(lambda: a.t() @ a)}
    line["gemm_nt_tflops_fp32_equivalent"] = flops / (line["gemm_nt_1024_ms"] * 1e-3) / 1e12
    line["matmul_nt_tflops"] = flops / (line["matmul_nt_1024_ms"] * 1e-3) / 1e12
    print(json.dumps(line), flush=True)
