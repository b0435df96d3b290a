"""A/B of the multi-step cluster kernel vs one launch per half-step (RBM_B200_MULTISTEP=0/1) on mid-size jobs."""
import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import divae_b200
torch.manual_seed(32)
for n, B, k in ((256, 1024, 20), (512, 2048, 100), (512, 4096, 100), (1024, 2048, 100), (1024, 4096, 100), (1024, 8192, 50)):
    rbm = divae_b200.RBM(n, n).cuda()
    s = divae_b200.PCD(batch_size=B, RBM=rbm, n_gibbs_sampling_steps=k, seed=1, variant="umma")
    for _ in range(3): s.block_gibbs_sampling()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.block_gibbs_sampling(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = statistics.median(ts)
    print(f"mode={os.environ.get('RBM_B200_MULTISTEP','auto')} {n}x{n} B={B} k={k}: {ms:.3f} ms/call, {ms*1e3/(2*k):.2f} us/half-step, {B*k*2*n/(ms*1e-3):.3e} unit-updates/s")
