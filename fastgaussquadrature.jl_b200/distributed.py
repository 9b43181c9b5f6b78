"""Optional collectives around the sharded, device-resident rules (one process per GPU,
``torch.distributed``: NCCL over NVLink on the GPUs, gloo in the CPU tests).

The hot path has NO collective: every node is a function of (n, index), so each rank writes its
own slice and nothing is exchanged.  What lives here is what the north star
lists as optional: an all-gather into a replicated array, and the all-reduce of the two moment
sums used as the sum-w / int x^2 check.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def allreduce_moments(sums: torch.Tensor) -> torch.Tensor:
    """Sum the per-rank (sum w, sum w x^2) pairs over all ranks (2 doubles on the wire)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return sums


def allgather_rule(shard):
    """Replicate the full rule on every rank: returns (x, w) of length n.  Each rank contributes
    its segments; segment extents are recomputed from (n, world, mode), so only payload travels."""
    from .api import half_index_partition, index_partition
    n, world, mode = shard.n, shard.world, shard.mode
    ref = shard.segments[0][2]
    x = torch.empty(n, dtype=torch.float64, device=ref.device)
    w = torch.empty(n, dtype=torch.float64, device=ref.device)
    if mode == "contiguous":
        extents = [[(lo, hi)] for lo, hi in index_partition(n, world)]
    else:
        m = (n + 1) // 2
        extents = []
        for k_lo, k_hi in half_index_partition(n, world):
            skip = 1 if (n % 2 and k_hi - 1 == m and k_hi > k_lo) else 0
            extents.append([(k_lo - 1, k_hi - 1), (n - k_hi + 1 + skip, n - k_lo + 1)])
    single = not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)
    for seg in range(len(extents[0])):
        for full, col in ((x, 2), (w, 3)):
            outs = [full[lo:hi] for lo, hi in (e[seg] for e in extents)]
            mine = shard.segments[seg][col].contiguous()
            if single:
                outs[0].copy_(mine)
            else:
                # uneven sizes: one broadcast per source rank (NCCL groups them over NVLink)
                for r, o in enumerate(outs):
                    if r == shard.rank:
                        o.copy_(mine)
                    dist.broadcast(o, src=r)
    return x, w
