"""Small driver for ncu captures of the shared-memory variant (BASELINE configs 2 and 3 at config,
and the steady-state case).  Usage: python benchmarks/smem_profile_target.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import divae_b200  # noqa: E402

torch.manual_seed(32)
for n, B, k in ((64, 100, 10), (256, 1024, 20), (64, 148 * 128 * 4, 200)):
    rbm = divae_b200.RBM(n, n).cuda()
    s = divae_b200.PCD(batch_size=B, RBM=rbm, n_gibbs_sampling_steps=k, seed=0x5EED, variant="smem")
    for _ in range(3):
        s.block_gibbs_sampling()
    torch.cuda.synchronize()
print("ok")
