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


def _gather_ranges(full: torch.Tensor, mine: torch.Tensor, ranges, rank: int, single: bool) -> int:
    """All-gather of one array: rank r contributes `mine` (= full[ranges[r][0]:ranges[r][1]] of the replica).
    ONE collective per array: straight into place when the partition is regular (equal lengths, rank-ascending and
    contiguous — the case for n = 1e9 on 1/2/4/8 ranks), otherwise through a padded staging buffer.  Returns the
    number of collectives issued."""
    if single:
        lo, hi = ranges[0]
        full[lo:hi].copy_(mine)
        return 0
    world = len(ranges)
    lens = [hi - lo for lo, hi in ranges]
    regular = len(set(lens)) == 1 and all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
    if regular and lens[0] > 0:
        dist.all_gather_into_tensor(full[ranges[0][0]:ranges[-1][1]], mine.contiguous())
        return 1
    mx = max(lens)
    if mx == 0:
        return 0
    padded = torch.zeros(mx, dtype=full.dtype, device=full.device)
    padded[:lens[rank]].copy_(mine)
    stage = torch.empty(world * mx, dtype=full.dtype, device=full.device)
    dist.all_gather_into_tensor(stage, padded)
    for r, (lo, hi) in enumerate(ranges):
        full[lo:hi].copy_(stage[r * mx:r * mx + (hi - lo)])
    return 1


def allgather_rule(shard, out=None, stats: dict | None = None):
    """Replicate the full rule on every rank: returns (x, w) of length n.

    One padded / in-place ``all_gather_into_tensor`` per array (NCCL over NVLink).  Mirror-pair shards ship only
    their LEFT segments — the right half of a replica is x[n-1-i] = -x[i], w[n-1-i] = w[i] (what the reference
    itself does, gausslegendre.jl:83-90) and is written locally by ``fgq_gausslegendre_mirror_device`` — so a
    replica costs n/2 * 16 * (world-1)/world bytes of inbound NVLink traffic instead of twice that.  Segment
    extents are recomputed from (n, world, mode): only payload travels."""
    from ._lib import check, lib
    from .api import half_index_partition, index_partition
    n, world, mode, rank = shard.n, shard.world, shard.mode, shard.rank
    ref = shard.segments[0][2]
    if out is None:
        x = torch.empty(n, dtype=torch.float64, device=ref.device)
        w = torch.empty(n, dtype=torch.float64, device=ref.device)
    else:
        x, w = out
    single = not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)
    ncoll = 0
    if mode == "contiguous":
        ranges = index_partition(n, world)
        ncoll += _gather_ranges(x, shard.segments[0][2], ranges, rank, single)
        ncoll += _gather_ranges(w, shard.segments[0][3], ranges, rank, single)
    else:
        ranges = [(k_lo - 1, k_hi - 1) for k_lo, k_hi in half_index_partition(n, world)]
        ncoll += _gather_ranges(x, shard.segments[0][2], ranges, rank, single)
        ncoll += _gather_ranges(w, shard.segments[0][3], ranges, rank, single)
        if x.is_cuda:
            check(lib().fgq_gausslegendre_mirror_device(n, x.data_ptr(), w.data_ptr(),
                                                        int(torch.cuda.current_stream().cuda_stream)))
        else:  # gloo / CPU tests of the host logic
            h = n // 2
            x[n - h:] = -x[:h].flip(0)
            w[n - h:] = w[:h].flip(0)
    if stats is not None:
        stats["collectives"] = ncoll
    return x, w
