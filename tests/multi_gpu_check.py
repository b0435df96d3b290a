"""Multi-GPU checks, launched by torchrun (one process per GPU, NCCL):
   sharded chains == slices of the single-GPU job; allreduced negative-phase statistics ==
   single-GPU statistics; sharded AIS log Z == single-GPU AIS log Z.
Run:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
          --master-port 29533 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import divae_b200  # noqa: E402
from divae_b200 import _ops, dist as bdist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    msg = run_checks(rank, world, dev)
    if rank == 0:
        print(msg)
    dist.destroy_process_group()


def run_checks(rank, world, dev):
    """All assertions of this file on an initialised process group (also run by bench.py before every N > 1 run)."""
    torch.manual_seed(32)
    total, k = 4096, 5
    for n, variant in ((256, "smem"), (512, "umma")):
        rbm = divae_b200.RBM(n, n)
        with torch.no_grad():
            rbm._weights.mul_(10.0)
        rbm = rbm.to(dev)
        # reference: the whole job on this GPU
        full = divae_b200.PCD(batch_size=total, RBM=rbm, n_gibbs_sampling_steps=k, seed=77, variant=variant)
        vf, hf = full.block_gibbs_sampling()
        Sf, svf, shf = _ops.negative_phase_stats(vf, hf)
        off, nloc = bdist.shard_chains(total, rank, world)
        shard = divae_b200.PCD(batch_size=nloc, RBM=rbm, n_gibbs_sampling_steps=k, seed=77, variant=variant, chain_offset=off)
        e, (S, sv, sh) = shard.negative_phase()
        torch.testing.assert_close(S, Sf, rtol=0, atol=1e-6)
        torch.testing.assert_close(sv, svf, rtol=0, atol=1e-6)
        torch.testing.assert_close(sh, shf, rtol=0, atol=1e-6)
        # default arguments (no chain_offset): the ranks' batches are laid out one after the other by negative_phase()
        auto = divae_b200.PCD(batch_size=nloc, RBM=rbm, n_gibbs_sampling_steps=k, seed=77, variant=variant)
        _, (S3, sv3, sh3) = auto.negative_phase()
        torch.testing.assert_close(S3, Sf, rtol=0, atol=1e-6)
        torch.testing.assert_close(sv3, svf, rtol=0, atol=1e-6)
        shard2 = divae_b200.PCD(batch_size=nloc, RBM=rbm, n_gibbs_sampling_steps=k, seed=77, variant=variant, chain_offset=off)
        vs, hs = shard2.block_gibbs_sampling()
        assert torch.equal(vs, vf[off:off + nloc]) and torch.equal(hs, hf[off:off + nloc])
        # the fused energy expectation backpropagates the allreduced statistics
        rbm.zero_grad()
        (-e).backward()
        torch.testing.assert_close(rbm._weights.grad, Sf, rtol=0, atol=1e-6)
    # shower generation sharded over the ranks: the gathered blocks equal the single-GPU generation
    from divae_b200.decoder import generate_samples

    class _V3(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self._activation_fct = torch.nn.ReLU()
            self._layers = torch.nn.ModuleList([torch.nn.Linear(129, 200), torch.nn.Linear(200, 300)])
            self._layers2 = torch.nn.ModuleList([torch.nn.Linear(300, 400), torch.nn.Linear(400, 504)])
            self._layers3 = torch.nn.ModuleList([torch.nn.Linear(300, 400), torch.nn.Linear(400, 504)])

    class _Model:
        pass

    torch.manual_seed(5)
    model = _Model()
    model.sampler = divae_b200.PCD(batch_size=128, RBM=divae_b200.RBM(64, 64).to(dev), n_gibbs_sampling_steps=10, seed=9)
    model.decoder = _V3().to(dev)
    single = _Model()
    single.sampler = divae_b200.PCD(batch_size=128, RBM=model.sampler._RBM, n_gibbs_sampling_steps=10, seed=9)
    single.decoder = model.decoder
    g1 = None                                        # a one-rank group = the unsharded job on this GPU
    for r in range(world):                           # (new_group is collective: every rank creates every group)
        g = dist.new_group([r])
        if r == rank:
            g1 = g
    e_full, s_full = generate_samples(single, num_samples=1024, true_energy=40.0, seed=3, group=g1)
    e_all, s_all = generate_samples(model, num_samples=1024, true_energy=40.0, seed=3, gather=True)
    assert s_all.shape == (1024, 504) and torch.equal(s_all, s_full) and torch.equal(e_all, e_full)
    # AIS: sharded estimate equals the unsharded one (same global chain ids)
    rbm = divae_b200.RBM(96, 64)
    with torch.no_grad():
        rbm._weights.mul_(20.0)
    rbm = rbm.to(dev)
    z_sharded = rbm.estimate_logZ(n_chains=1024, n_betas=100, seed=5)
    dist.barrier()
    # single-GPU value: run with no group by temporarily computing all chains locally
    import ctypes
    from divae_b200 import _lib
    lib = _lib.load()
    p, keep = rbm._pack.pack(rbm, need_bf16=True)
    betas = np.linspace(0.0, 1.0, 101)
    logw = torch.empty(1024, dtype=torch.float64, device=dev)
    nbytes = lib.rbm_b200_ais_workspace(96, 64, 1024)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    # same weight representation as estimate_logZ's default (W rounded to bf16 in the library); the legacy entry
    # rbm_b200_ais_run without a W_bf16 pointer would anneal with split (fp32-grade) weights: 0.02 nats away here
    _lib.check(lib.rbm_b200_ais_run_ex(ctypes.byref(p), 1024, betas.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 101, 5, 0,
                                       _ops._ptr(logw), _lib.VARIANTS["umma"], _ops._ptr(ws), nbytes, _ops._stream()))
    a = rbm._visible_bias.detach().double()
    z_single = (64 * np.log(2.0) + torch.nn.functional.softplus(a).sum() + torch.logsumexp(logw, 0) - np.log(1024)).item()
    assert abs(z_sharded - z_single) < 1e-9, (z_sharded, z_single)
    dist.barrier()
    return ("multi-GPU checks ok on %d GPUs: shards == slices, allreduced stats == single-GPU stats, "
            "sharded generation == single-GPU generation, AIS log Z %.6f" % (world, z_sharded))


if __name__ == "__main__":
    main()
