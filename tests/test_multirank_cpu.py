"""world_size-2 gloo test (CPU) of the N>1 host logic: each rank takes its shard from the
partition functions the GPU path uses, the shards tile the rule with no exchange, and the only
collectives are the optional all-gather and the 2-double moment all-reduce.  The oracle stands in
for the kernels here (tests may use it; the product never does)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, mode, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fgq_b200
    import oracle
    from fgq_b200 import distributed as fd
    xf, wf = oracle.gausslegendre(n, twin=True)
    if mode == "contiguous":
        lo, hi = fgq_b200.index_partition(n, world)[rank]
        segs = [(lo, hi, torch.from_numpy(xf[lo:hi].copy()), torch.from_numpy(wf[lo:hi].copy()))]
    else:
        k_lo, k_hi = fgq_b200.half_index_partition(n, world)[rank]
        xh, wh = oracle.legendre_asy_half(n, k_lo, k_hi - 1, twin=True)
        m = (n + 1) // 2
        skip = 1 if (n % 2 and k_hi - 1 == m) else 0
        xl = -xh
        if skip:
            xl[-1] = 0.0
        segs = [(k_lo - 1, k_hi - 1, torch.from_numpy(xl), torch.from_numpy(wh.copy())),
                (n - k_hi + 1 + skip, n - k_lo + 1,
                 torch.from_numpy(xh[::-1][skip:].copy()), torch.from_numpy(wh[::-1][skip:].copy()))]
    shard = fgq_b200.LegendreShard(n, rank, world, mode, segs)
    x, w = fd.allgather_rule(shard)
    sums = fd.allreduce_moments(torch.tensor(
        [sum(float(s[3].sum()) for s in segs), sum(float((s[3] * s[2] * s[2]).sum()) for s in segs)],
        dtype=torch.float64))
    ok = np.array_equal(x.numpy(), xf) and np.array_equal(w.numpy(), wf)
    ok = ok and abs(float(sums[0]) - 2) < 1e-12 and abs(float(sums[1]) - 2 / 3) < 1e-12
    open(os.path.join(out_dir, f"rank{rank}.ok"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["pair", "contiguous"])
@pytest.mark.parametrize("n,world", [(1000, 2), (100001, 2), (100003, 3)])
def test_multi_rank_sharding_and_collectives(tmp_path, n, world, mode):
    """world 2: regular partitions (one in-place all-gather per array); world 3 with an odd n: unequal shards
    (padded all-gather); pair mode ships left segments only and mirrors locally."""
    mp.spawn(_worker, args=(world, _free_port(), n, mode, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"rank{r}.ok").read() == "1"
