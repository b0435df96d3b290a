"""cProfile of the Python side of PCD.block_gibbs_sampling() (config 2: the call is host-bound)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import divae_b200  # noqa: E402

rbm = divae_b200.RBM(64, 64).cuda()
s = divae_b200.PCD(batch_size=100, RBM=rbm, n_gibbs_sampling_steps=10, seed=1)
for _ in range(50):
    s.block_gibbs_sampling()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000):
    s.block_gibbs_sampling()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("issue %.2f us per call" % ((t1 - t0) / 2000 * 1e6))
pr = cProfile.Profile()
pr.enable()
for _ in range(2000):
    s.block_gibbs_sampling()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
