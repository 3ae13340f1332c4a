"""world_size-2 gloo test of the N>1 host logic (partition, per-rank generation, reductions):
the data path has no collective, so this is all that multi-GPU adds."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from piplib_b200 import dist as pdist
from workloads import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = pdist.problem_range(rank, world, B)
    dom, ctx = synth.generate("loopnest8x12p2", n, seed=9, first=first)
    times, counts = pdist.reduce_stats([10.0 + rank, 5.0 - rank], [float(n), float(dom[:, :, -1].sum())])
    q.put((rank, first, n, int(dom[:, :, -1].sum()), times, counts, dom[0].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_and_reduction():
    world, B = 2, 3000
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # ranges are disjoint and cover [0, world*B)
    assert [(r[1], r[2]) for r in res] == [(0, B), (B, B)]
    # a rank's slice is exactly the slice of one big generation (counter-based generator)
    big, _ = synth.generate("loopnest8x12p2", world * B, seed=9, first=0)
    for rank, first, n, s, times, counts, row0 in res:
        assert s == int(big[first:first + n, :, -1].sum())
        assert row0 == big[first].tolist()
        assert times == [11.0, 5.0]                       # max over ranks
        assert counts[0] == world * B                     # sum over ranks
        assert counts[1] == float(big[:, :, -1].sum())


def test_split_range():
    parts = pdist.split_range(10, 103, 8)
    assert parts[0][0] == 10 and sum(n for _, n in parts) == 103
    assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(7))
    assert max(n for _, n in parts) - min(n for _, n in parts) <= 1
