"""Multi-GPU plumbing: chains shard across ranks, two small collectives.

One process per GPU (torch.distributed, NCCL on GPUs, gloo in CPU tests).  Chains are
independent given (W, a, c), so the data path has NO collective; because every draw is
keyed by the GLOBAL chain id, a sharded job draws exactly what the unsharded job would.
(Kernel choice: `variant="auto"` looks at the LOCAL chain count.  Every choice it can make for a given prior multiplies the
same bf16-rounded W and follows the same draw contract, so shards and the full job agree except on chains that met a
draw within 1e-6 of its threshold - the kernels sum in different orders; pass an explicit `variant=` on every rank when
the shards must be bit-identical slices, as tests/multi_gpu_check.py does.)
Collectives exist only for
  * the sum-allreduce of the fused negative-phase statistics (one flat buffer), and
  * the log-sum-exp reduction of AIS log-weights (max, then sum of exp).
"""
import math

import torch
import torch.distributed as dist


def shard_chains(total_chains, rank, world_size):
    """Contiguous block partition: returns (chain_offset, n_local). Remainder goes to low ranks."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(int(total_chains), int(world_size))
    n_local = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, n_local


def allreduce_stats(S_vh, S_v, S_h, group=None):
    """Sum-allreduce (S_vh, S_v, S_h) as ONE flat buffer (latency-bound: <= 4.2 MiB at n=1024)."""
    flat = torch.cat([S_vh.reshape(-1), S_v.reshape(-1), S_h.reshape(-1)])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    n0, n1 = S_vh.numel(), S_v.numel()
    return flat[:n0].view_as(S_vh), flat[n0:n0 + n1].view_as(S_v), flat[n0 + n1:].view_as(S_h)


def logsumexp_allreduce(local_logw, group=None):
    """log sum_i exp(logw_i) over all ranks' shards: allreduce(max) then allreduce(sum)."""
    if local_logw.numel() == 0:
        m_local = torch.full((), -math.inf, dtype=torch.float64, device=local_logw.device)
    else:
        m_local = local_logw.max().double()
    m = m_local.clone()
    if dist.is_initialized():
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    s = torch.exp(local_logw.double() - m).sum() if local_logw.numel() else torch.zeros((), dtype=torch.float64,
                                                                                         device=local_logw.device)
    if dist.is_initialized():
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    return m + torch.log(s)
