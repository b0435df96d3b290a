"""Worker of tests/test_gpu_chain.py: runs block Gibbs / AIS under the tile overrides of the environment it was started
with (RBM_B200_CHAIN*, read once per process) and stores the results in an .npz file."""
import sys

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import divae_b200  # noqa: E402


def main():
    out = sys.argv[1]
    res = {}
    rng = np.random.default_rng(123)
    for (nV, nH, B, k, scale) in ((1024, 1024, 700, 3, 0.03), (320, 192, 1500, 4, 0.1), (512, 512, 300, 5, 0.05),
                                  (64, 64, 128, 2, 0.3), (130, 700, 260, 3, 0.05),
                                  (256, 256, 4500, 2, 0.1)):      # 36 row blocks: more than one round of resident clusters
        W = (rng.standard_normal((nV, nH)) * scale).astype(np.float32)
        W = torch.from_numpy(W).to(torch.bfloat16).float().numpy()
        a = (0.5 + 0.2 * rng.standard_normal(nV)).astype(np.float32)
        c = (0.2 * rng.standard_normal(nH)).astype(np.float32)
        rbm = divae_b200.RBM(nV, nH)
        with torch.no_grad():
            rbm._weights.copy_(torch.from_numpy(W)); rbm._visible_bias.copy_(torch.from_numpy(a))
            rbm._hidden_bias.copy_(torch.from_numpy(c))
        rbm = rbm.to("cuda:0")
        s = divae_b200.PCD(batch_size=B, RBM=rbm, n_gibbs_sampling_steps=k, seed=0x5EED, variant="umma", chain_offset=77)
        v, h = s.block_gibbs_sampling()
        key = "%dx%d" % (nV, nH)
        res["v_" + key] = np.packbits(v.cpu().numpy().astype(np.uint8), axis=1)
        res["h_" + key] = np.packbits(h.cpu().numpy().astype(np.uint8), axis=1)
        if (nV, nH) == (320, 192):
            res["W"], res["a"], res["c"] = W, a, c
            z, logw = rbm.estimate_logZ(n_chains=600, n_betas=60, seed=11, return_logw=True)
            res["logw"] = logw.cpu().numpy()
    torch.cuda.synchronize()
    np.savez(out, **res)


if __name__ == "__main__":
    main()
