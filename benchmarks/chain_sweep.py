"""Tile-shape sweep of the dataflow chain kernel (umma_chain.cuh) against the per-half-step launches.

  python benchmarks/chain_sweep.py [--n 1024] [--k 200] [--chains 8192,16384,65536] [--out gpurun_out/chain_sweep.jsonl]

Every configuration runs in its own process (the tile overrides are read once per process):
RBM_B200_CHAIN=0|1, RBM_B200_CHAIN_BN=64|128|256, RBM_B200_CHAIN_PAIRS=0|1.  Prints one JSON line per
configuration: microseconds per half-step and the tensor throughput it corresponds to.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import json, os, sys, torch
sys.path.insert(0, %(root)r)
import divae_b200
n, B, k, reps = %(n)d, %(B)d, %(k)d, %(reps)d
torch.manual_seed(32)
rbm = divae_b200.RBM(n, n).to("cuda:0")
s = divae_b200.PCD(batch_size=B, RBM=rbm, n_gibbs_sampling_steps=k, seed=0x5EED, variant="%(variant)s")
for _ in range(2):
    s.block_gibbs_sampling()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    v, h = s.block_gibbs_sampling()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
us = ms * 1e3 / (2 * k)
print(json.dumps({"n": n, "chains": B, "k": k, "variant": "%(variant)s", "ms_per_call": ms, "us_per_half_step": us,
                  "tflops": 2.0 * B * n * n / (us * 1e-6) / 1e12, "unit_updates_per_s": B * k * 2 * n / (ms * 1e-3),
                  "mean_v": float(v.mean())}))
"""

AIS_WORKER = r"""
import json, os, sys, torch, numpy as np
sys.path.insert(0, %(root)r)
import divae_b200
n, M, K, reps = %(n)d, %(B)d, %(k)d, %(reps)d
torch.manual_seed(32)
rbm = divae_b200.RBM(n, n)
with torch.no_grad():
    rbm._weights.mul_(5.0)
rbm = rbm.to("cuda:0")
betas = np.linspace(0.0, 1.0, K + 1)
z = rbm.estimate_logZ(n_chains=M, betas=betas, seed=0x5EED, variant="%(variant)s")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    z = rbm.estimate_logZ(n_chains=M, betas=betas, seed=0x5EED, variant="%(variant)s")
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
us = ms * 1e3 / (2 * K - 1)
print(json.dumps({"ais": True, "n": n, "chains": M, "betas": K, "variant": "%(variant)s", "ms_per_call": ms, "us_per_half_step": us,
                  "tflops": 2.0 * M * n * n / (us * 1e-6) / 1e12, "unit_updates_per_s": M * K * 2 * n / (ms * 1e-3),
                  "log_Z": z}))
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--k", type=int, default=200)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--chains", default="8192,16384,65536")
    ap.add_argument("--variant", default="umma")
    ap.add_argument("--configs", default="launches,bn256p,bn256,bn128,bn64,auto")
    ap.add_argument("--out", default=None)
    ap.add_argument("--ais", action="store_true", help="time RBM.estimate_logZ (k = beta steps) instead of block Gibbs sampling")
    args = ap.parse_args()
    envs = {
        "launches": {"RBM_B200_CHAIN": "0"},
        "bn256p": {"RBM_B200_CHAIN_BN": "256", "RBM_B200_CHAIN_PAIRS": "1"},
        "bn256": {"RBM_B200_CHAIN_BN": "256", "RBM_B200_CHAIN_PAIRS": "0"},
        "bn128": {"RBM_B200_CHAIN_BN": "128"},
        "bn64": {"RBM_B200_CHAIN_BN": "64"},
        "auto": {},
        "sb8k": {"RBM_B200_CHAIN_SB": "8192"},
        "sb16k": {"RBM_B200_CHAIN_SB": "16384"},
        "sb32k": {"RBM_B200_CHAIN_SB": "32768"},
        "sb1": {"RBM_B200_CHAIN_SB": "-1"},
        "opts0": {"RBM_B200_CHAIN_OPTS": "0"},
        "opts1": {"RBM_B200_CHAIN_OPTS": "1"},
        "opts4": {"RBM_B200_CHAIN_OPTS": "4"},
        "opts5": {"RBM_B200_CHAIN_OPTS": "5"},
        "opts6": {"RBM_B200_CHAIN_OPTS": "6"},
        "opts7": {"RBM_B200_CHAIN_OPTS": "7"},
        "grid64": {"RBM_B200_CHAIN_GRID": "64"},
        "grid70": {"RBM_B200_CHAIN_GRID": "70"},
    }
    out = open(args.out, "a") if args.out else None
    for B in [int(x) for x in args.chains.split(",")]:
        for name in args.configs.split(","):
            env = dict(os.environ)
            for key in ("RBM_B200_CHAIN", "RBM_B200_CHAIN_BN", "RBM_B200_CHAIN_PAIRS", "RBM_B200_CHAIN_SB", "RBM_B200_CHAIN_GRID", "RBM_B200_CHAIN_OPTS"):
                env.pop(key, None)
            if name.startswith("opts") and name not in envs:
                envs[name] = {"RBM_B200_CHAIN_OPTS": name[4:]}
            env.update(envs[name])
            code = (AIS_WORKER if args.ais else WORKER) % dict(root=ROOT, n=args.n, B=B, k=args.k, reps=args.reps, variant=args.variant)
            try:
                r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
            except subprocess.TimeoutExpired:
                line = json.dumps({"config": name, "chains": B, "error": "timeout"})
                print(line, flush=True)
                if out:
                    out.write(line + "\n"); out.flush()
                continue
            if r.returncode != 0:
                line = json.dumps({"config": name, "chains": B, "error": r.stderr[-400:]})
            else:
                d = json.loads(r.stdout.strip().splitlines()[-1])
                d["config"] = name
                line = json.dumps(d)
            print(line, flush=True)
            if out:
                out.write(line + "\n"); out.flush()


if __name__ == "__main__":
    main()
