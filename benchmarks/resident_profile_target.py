"""Small driver for ncu captures of the resident kernel (umma_resident.cuh): an AIS shard of BASELINE configs[4]
(512 x 512, 2048 chains = one row block per cluster) and a ping-pong case (16384 chains), short schedules.
Usage: python benchmarks/resident_profile_target.py [chains] [betas]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import divae_b200  # noqa: E402

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
K = int(sys.argv[2]) if len(sys.argv) > 2 else 300
torch.manual_seed(32)
rbm = divae_b200.RBM(512, 512)
with torch.no_grad():
    rbm._weights.mul_(5.0)
rbm = rbm.cuda()
betas = np.linspace(0.0, 1.0, K + 1)
for _ in range(2):
    z = rbm.estimate_logZ(n_chains=chains, betas=betas, seed=0x5EED)
torch.cuda.synchronize()
print("ok", z)
