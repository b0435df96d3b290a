import os, sys, torch
sys.path.insert(0, "/root/repo")
import divae_b200
torch.manual_seed(0)
rbm = divae_b200.RBM(64, 64).cuda()
s = divae_b200.PCD(batch_size=128, RBM=rbm, n_gibbs_sampling_steps=40, seed=1)
for _ in range(20): s.negative_phase()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): s.negative_phase()
e1.record(); torch.cuda.synchronize()
print("negative_phase() PCD(128,k=40) 64x64:", e0.elapsed_time(e1) / 200 * 1e3, "us/call")
