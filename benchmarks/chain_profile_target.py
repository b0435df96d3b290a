"""Small driver for ncu captures of the dataflow chain kernel (umma_chain.cuh): one shard of BASELINE configs[3] at
8 GPUs (RBM 1024 x 1024, 8,192 chains), k Gibbs steps per call.  Usage: python benchmarks/chain_profile_target.py [chains] [k]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import divae_b200  # noqa: E402

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
k = int(sys.argv[2]) if len(sys.argv) > 2 else 50
torch.manual_seed(32)
rbm = divae_b200.RBM(1024, 1024).cuda()
s = divae_b200.PCD(batch_size=chains, RBM=rbm, n_gibbs_sampling_steps=k, seed=0x5EED)
for _ in range(3):
    v, h = s.block_gibbs_sampling()
torch.cuda.synchronize()
print("ok", float(v.mean()))
