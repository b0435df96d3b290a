"""BASELINE config 5 under torchrun: AIS log Z of a 512x512 RBM, 16,384 chains x 10,000 betas sharded
over the ranks, NCCL log-sum-exp of the log-weights.  Prints one JSON line on rank 0.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29571 \
      benchmarks/ais_multi_gpu.py [--betas 10000] [--chains 16384]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import divae_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--betas", type=int, default=10000)
    ap.add_argument("--chains", type=int, default=16384)
    ap.add_argument("--n", type=int, default=512)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(32)
    rbm = divae_b200.RBM(args.n, args.n)
    with torch.no_grad():
        rbm._weights.mul_(5.0)       # trained-like scale so that the estimate is non-trivial
    rbm = rbm.to(dev)
    rbm.estimate_logZ(n_chains=args.chains, n_betas=max(10, args.betas // 100), seed=1)   # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    logZ = rbm.estimate_logZ(n_chains=args.chains, n_betas=args.betas, seed=1)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        units = args.chains * args.betas * 2 * args.n
        print(json.dumps({"config": "5: AIS log Z %dx%d, %d chains x %d betas" % (args.n, args.n, args.chains, args.betas),
                          "n_gpus": world, "ms": ms.item(), "unit_updates_per_s": units / (ms.item() * 1e-3), "logZ": logZ}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
