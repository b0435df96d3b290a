"""Latency of one CD-1 update at BASELINE config 1 (vanilla RBM 784 x 128, batch 128): the fused cooperative
kernel (default) against the multi-launch path (RBM_B200_CD_FUSED=0)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from divae_b200.samplers.contrastiveDivergence import ContrastiveDivergence  # noqa: E402

torch.manual_seed(0)
data = (torch.rand(128, 784, device="cuda") > 0.5).float()
cd = ContrastiveDivergence(784, 128, 1e-3, 0.5, 1, 1e-4, device="cuda", seed=3)
for _ in range(20):
    cd.contrastive_divergence(data)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(300):
    loss = cd.contrastive_divergence(data)
e1.record()
torch.cuda.synchronize()
print("CD-1 784x128 batch 128, RBM_B200_CD_FUSED=%s: %.1f us per update, loss %.2f"
      % (os.environ.get("RBM_B200_CD_FUSED", "auto"), e0.elapsed_time(e1) / 300 * 1e3, loss.item()))
