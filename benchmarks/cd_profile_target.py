"""Driver for ncu launch lists of one CD-1 update (BASELINE config 1)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from divae_b200.samplers.contrastiveDivergence import ContrastiveDivergence
cd = ContrastiveDivergence(784, 128, 1e-3, 0.5, 1, 1e-4, device="cuda:0", seed=1)
data = (torch.rand(128, 784, device="cuda:0") > 0.5).float()
for _ in range(5):
    cd.contrastive_divergence(data)
torch.cuda.synchronize()
print("ok")
