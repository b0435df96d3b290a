"""torch-CPU port of the reference sampler, op for op — the TIMED CPU baseline.

TEST/BENCH INFRASTRUCTURE (see oracle/__init__.py).  The reference is pure
PyTorch and cannot travel to the GPU box (/root/reference does not exist
there), so `bench.py`'s `cpu_baseline` and `--impl reference` legs time this
port instead: the same ATen calls in the same order as
models/samplers/pcd.py:23-69 (matmul, add, sigmoid, rand, ge, float), under
torch.no_grad like the benchmark plan in BASELINE.md §4.  oracle/gen_golden.py
checks, in the build container, that with the same torch seed it returns
bit-identical tensors to the reference's own PCD.
"""
import torch


def pcd_init_state(batch_size, n_visible):
    """models/samplers/pcd.py:16-17."""
    return (torch.rand(batch_size, n_visible) >= torch.rand(batch_size, n_visible)).float()


def hidden_samples(v, W, c):
    """models/samplers/pcd.py:32-35."""
    hidden_probs = torch.sigmoid(torch.matmul(v, W) + c)
    return (hidden_probs >= torch.rand(hidden_probs.size(), device=hidden_probs.device)).float()


def visible_samples(h, W, a):
    """models/samplers/pcd.py:46-49."""
    visible_probs = torch.sigmoid(torch.matmul(h, W.t()) + a)
    return (visible_probs >= torch.rand(visible_probs.size(), device=visible_probs.device)).float()


def block_gibbs_sampling(mc_state, W, a, c, k):
    """models/samplers/pcd.py:51-69 (returns (v, h, new_mc_state))."""
    v = mc_state.to(a.device)
    h = None
    for _ in range(k):
        h = hidden_samples(v, W, c)
        v = visible_samples(h, W, a)
    new_state = pcd_init_state(mc_state.shape[0], a.size(0))
    return v, h, new_state


def cd1_step(v0, W, a, c, k=1):
    """sandbox/rbm_standalone.py:66-73,75-91 statistics of one CD-k call (timed for config 1)."""
    ph0 = torch.sigmoid(torch.matmul(v0, W) + c)
    h = (ph0 >= torch.rand(W.shape[1])).float()
    pos = torch.matmul(v0.t(), h)
    for _ in range(k):
        pv = torch.sigmoid(torch.matmul(h, W.t()) + a)
        ph = torch.sigmoid(torch.matmul(pv, W) + c)
        h = (ph >= torch.rand(W.shape[1])).float()
    neg = torch.matmul(pv.t(), ph)
    return pos - neg, torch.sum(v0 - pv, dim=0), torch.sum(ph0 - ph, dim=0)
