"""Arithmetic error of the AIS log-weights: GPU (fp32, MUFU approximations) against the float64 CPU restatement
(oracle/rbm_oracle.py:ais_logZ) on the SAME draws.  Chains that never meet a near-threshold draw follow the oracle's
trajectory exactly, so for them the difference is pure arithmetic error; its MEAN is the bias an estimate inherits.

  python benchmarks/ais_weight_accuracy.py n chains betas scale --make-ref ref.npz     (CPU: the oracle's log-weights, minutes)
  python benchmarks/ais_weight_accuracy.py n chains betas scale --ref ref.npz          (GPU box: compare)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import divae_b200  # noqa: E402
from oracle import rbm_oracle as orc  # noqa: E402  (benchmark-side checker, never on the product path)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
M = int(sys.argv[2]) if len(sys.argv) > 2 else 256
K = int(sys.argv[3]) if len(sys.argv) > 3 else 400
scale = float(sys.argv[4]) if len(sys.argv) > 4 else 0.05
rng = np.random.default_rng(5)
W = (rng.standard_normal((n, n)) * scale).astype(np.float32)
W = torch.from_numpy(W).to(torch.bfloat16).float().numpy()
a = (0.2 * rng.standard_normal(n)).astype(np.float32)
c = (0.2 * rng.standard_normal(n)).astype(np.float32)
betas = np.linspace(0.0, 1.0, K + 1)
if "--make-ref" in sys.argv:
    ref, ref_logw, _ = orc.ais_logZ(W, a, c, M, betas, 11, return_logw=True)
    np.savez(sys.argv[sys.argv.index("--make-ref") + 1], ref=ref, ref_logw=ref_logw)
    print("oracle log Z", ref)
    sys.exit(0)
rbm = divae_b200.RBM(n, n)
with torch.no_grad():
    rbm._weights.copy_(torch.from_numpy(W)); rbm._visible_bias.copy_(torch.from_numpy(a)); rbm._hidden_bias.copy_(torch.from_numpy(c))
rbm = rbm.to("cuda:0")
est, logw = rbm.estimate_logZ(n_chains=M, betas=betas, seed=11, return_logw=True)
if "--ref" in sys.argv:
    z = np.load(sys.argv[sys.argv.index("--ref") + 1]); ref, ref_logw = float(z["ref"]), z["ref_logw"]
else:
    ref, ref_logw, _ = orc.ais_logZ(W, a, c, M, betas, 11, return_logw=True)
d = logw.cpu().numpy() - ref_logw
clean = np.abs(d) < 0.02
print(json.dumps({"n": n, "chains": M, "betas": K, "scale": scale, "logZ_gpu": est, "logZ_oracle": ref, "clean_chains": int(clean.sum()),
                  "mean_signed_diff": float(d[clean].mean()), "median_signed_diff": float(np.median(d[clean])),
                  "rms_diff": float(np.sqrt((d[clean] ** 2).mean())), "mean_logw": float(ref_logw.mean()),
                  "defs": os.environ.get("RBM_B200_NVCC_DEFS", "")}))
