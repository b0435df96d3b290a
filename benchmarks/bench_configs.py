#!/usr/bin/env python
"""All five BASELINE.json configs on ONE GPU: at-config latency and unit-updates/s, plus the
steady-state rate of the shared-memory variant (same prior, many chains).  One JSON line per row.

  python benchmarks/bench_configs.py [--quick]

Timing: CUDA events on the launching stream around `reps` back-to-back calls after 3 warm-ups;
median and best reported.  Not the contract benchmark (that is bench.py at the repo root).
"""
import argparse
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import divae_b200  # noqa: E402
from divae_b200.samplers.contrastiveDivergence import ContrastiveDivergence  # noqa: E402

DEV = "cuda:0"


def timeit(fn, reps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    return statistics.median(times), min(times)


def pcd_row(name, n, chains, k, variant, reps, scale=0.01):
    torch.manual_seed(32)
    rbm = divae_b200.RBM(n, n)
    with torch.no_grad():
        rbm._weights.mul_(scale / 0.01)
    rbm = rbm.to(DEV)
    s = divae_b200.PCD(batch_size=chains, RBM=rbm, n_gibbs_sampling_steps=k, seed=0x5EED, variant=variant)
    med, best = timeit(s.block_gibbs_sampling, reps)
    units = chains * k * 2 * n
    return {"config": name, "rbm": "%dx%d" % (n, n), "chains": chains, "gibbs_steps": k, "variant": variant,
            "ms_per_call_median": med, "ms_per_call_best": best, "unit_updates_per_s": units / (med * 1e-3),
            "flops_per_call": 4.0 * chains * k * n * n, "tflops": 4.0 * chains * k * n * n / (med * 1e-3) / 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    reps = 5 if args.quick else 20
    rows = []
    # config 1: vanilla RBM 784x128, CD-1, batch 128 (sandbox/rbm_standalone.py semantics)
    cd = ContrastiveDivergence(784, 128, 1e-3, 0.5, 1, 1e-4, device=DEV, seed=1)
    data = (torch.rand(128, 784, device=DEV) > 0.5).float()
    med, best = timeit(lambda: cd.contrastive_divergence(data), reps)
    rows.append({"config": "1: vanilla RBM 784x128 CD-1 batch 128", "ms_per_iter_median": med, "ms_per_iter_best": best,
                 "unit_updates_per_s": 128 * (128 + 1 * (784 + 128)) / (med * 1e-3), "variant": "general (fp32)"})
    # config 2 / 3 at config, both small-prior variants
    for variant in ("smem", "umma"):
        rows.append(pcd_row("2: DiVAE MNIST-shape PCD 100 chains x 10 steps", 64, 100, 10, variant, reps))
        rows.append(pcd_row("3: DiVAE++ calo PCD 1024 chains x 20 steps", 256, 1024, 20, variant, reps))
    rows.append(pcd_row("ref default: PCD(128, k=40) on the 64x64 prior (dvaepp.py:57)", 64, 128, 40, "smem", reps))
    # steady state of the small priors (chains >= 148*128, large k)
    for variant in ("smem", "umma"):
        rows.append(pcd_row("2 steady: 64x64, 148*128*4 chains x 200 steps", 64, 148 * 128 * 4, 200, variant, max(3, reps // 4)))
        rows.append(pcd_row("3 steady: 256x256, 148*128*4 chains x 100 steps", 256, 148 * 128 * 4, 100, variant, max(3, reps // 4)))
    # config 4 (reduced k here; bench.py runs the full one)
    rows.append(pcd_row("4: large prior 1024x1024, 65536 chains x 100 steps", 1024, 65536, 100, "umma", 3))
    # config 5: AIS 512x512, 16384 chains x 10000 betas on one GPU (8 GPUs shard the chains)
    torch.manual_seed(32)
    rbm = divae_b200.RBM(512, 512).to(DEV)
    nb = 1000 if args.quick else 10000
    med, best = timeit(lambda: rbm.estimate_logZ(n_chains=16384, n_betas=nb, seed=1), 1 if not args.quick else 2, warmup=1)
    rows.append({"config": "5: AIS log Z 512x512, 16384 chains x %d betas (1 GPU)" % nb, "ms_per_estimate": med,
                 "unit_updates_per_s": 16384 * nb * 1024 / (med * 1e-3), "logZ": rbm.get_logZ_value(),
                 "tflops": 4.0 * 16384 * nb * 512 * 512 / (med * 1e-3) / 1e12})
    for r in rows:
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
