"""ncu target: two PCD.negative_phase() calls on the 1024x1024 prior with 65,536 chains, k = 1."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import divae_b200  # noqa: E402

rbm = divae_b200.RBM(1024, 1024).cuda()
s = divae_b200.PCD(batch_size=65536, RBM=rbm, n_gibbs_sampling_steps=1, seed=1)
for _ in range(2):
    e, stats = s.negative_phase()
torch.cuda.synchronize()
print(float(e))
