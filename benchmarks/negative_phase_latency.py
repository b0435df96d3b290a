"""Latency of PCD.negative_phase() (sampler + fused <vh>, <v>, <h> statistics + energy expectation).

  python benchmarks/negative_phase_latency.py            # reference-default PCD(128, k=40) on 64x64
  python benchmarks/negative_phase_latency.py large      # 1024x1024, 65,536 chains: statistics on the tensor cores

`large` reports the call at k = 1 and k = 2 so the statistics part (everything except the 2k half-steps)
can be read off: t(k=1) - [t(k=2) - t(k=1)].
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import divae_b200  # noqa: E402


def time_call(fn, warmup, iters):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def main():
    torch.manual_seed(0)
    if len(sys.argv) > 1 and sys.argv[1] == "large":
        n, B = 1024, 65536
        rbm = divae_b200.RBM(n, n).cuda()
        out = {"shape": "%dx%d" % (n, n), "chains": B}
        for k in (1, 2):
            s = divae_b200.PCD(batch_size=B, RBM=rbm, n_gibbs_sampling_steps=k, seed=1)
            out["negative_phase_k%d_us" % k] = time_call(s.negative_phase, 3, 10)
            out["block_gibbs_k%d_us" % k] = time_call(s.block_gibbs_sampling, 3, 10)
        out["statistics_us"] = out["negative_phase_k1_us"] - out["block_gibbs_k1_us"]
        out["flops_vh"] = 2.0 * B * n * n
        out["vh_tflops_upper_bound"] = out["flops_vh"] / (out["statistics_us"] * 1e-6) / 1e12
        print(json.dumps(out))
        return
    rbm = divae_b200.RBM(64, 64).cuda()
    s = divae_b200.PCD(batch_size=128, RBM=rbm, n_gibbs_sampling_steps=40, seed=1)
    print("negative_phase() PCD(128,k=40) 64x64:", time_call(s.negative_phase, 20, 200), "us/call")


if __name__ == "__main__":
    main()
