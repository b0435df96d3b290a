"""ncu target: three shower-decoder passes (65,536 showers, calorimeter shape) through rbm_b200_decoder_forward."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchmarks.decoder_bench import V3  # noqa: E402
from divae_b200.decoder import DecoderRunner  # noqa: E402

torch.manual_seed(0)
B = 65536
m = V3([(129, 200), (200, 300), (300, 400), (400, 504)]).cuda()
r = DecoderRunner(m)
x = torch.cat([(torch.rand(B, 128, device="cuda") < 0.5).float(), torch.rand(B, 1, device="cuda") * 100], 1)
for _ in range(3):
    s = r.shower(x, seed=1)
torch.cuda.synchronize()
print(float(s.mean()))
