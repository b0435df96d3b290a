import time, torch, sys
sys.path.insert(0, '/root/repo')
import divae_b200
torch.manual_seed(0)
for n, B, k in ((64, 100, 10), (64, 128, 40), (256, 1024, 20)):
    rbm = divae_b200.RBM(n, n).cuda()
    s = divae_b200.PCD(batch_size=B, RBM=rbm, n_gibbs_sampling_steps=k, seed=1)
    for _ in range(20): s.block_gibbs_sampling()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(500): s.block_gibbs_sampling()
    t_issue = (time.perf_counter() - t0) / 500
    torch.cuda.synchronize()
    t_total = (time.perf_counter() - t0) / 500
    # pure GPU time via events around 100 calls
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): s.block_gibbs_sampling()
    e1.record(); torch.cuda.synchronize()
    print(f"{n}x{n} B={B} k={k}: host issue {t_issue*1e6:.1f} us/call, wall {t_total*1e6:.1f} us/call, event {e0.elapsed_time(e1)/200*1e3:.1f} us/call")
