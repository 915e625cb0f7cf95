"""N>1 host logic on CPU: two gloo ranks shard frames with no overlap and agree on the aggregate
(sum of frames / max elapsed), as bench.py does with nccl on GPUs."""
import os
import socket

import pytest

from flame_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = shard.frames_for_rank(rank, world, 4)
    elapsed = 0.5 + 0.25 * rank  # rank 1 is slower
    agg = shard.aggregate_throughput(len(frames), elapsed)
    tmax = shard.max_over_ranks(elapsed)
    dist.barrier()
    q.put((rank, frames, agg, tmax))
    dist.destroy_process_group()


def test_two_rank_sharding_and_aggregation():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, a0, t0), (r1, f1, a1, t1) = res
    assert f0 == [0, 1, 2, 3] and f1 == [4, 5, 6, 7] and not set(f0) & set(f1)
    assert t0 == t1 == 0.75 and a0 == a1 == pytest.approx(8 / 0.75)
    assert all(shard.frame_owner(f, 2, 4) == 0 for f in f0) and all(shard.frame_owner(f, 2, 4) == 1 for f in f1)


def _gather_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = torch.full((3, 4), float(rank + 1))
    m[0, 0] = float("nan")
    v = torch.arange(5 + 2 * rank, dtype=torch.float32) + 100 * rank  # different vertex counts per rank
    maps, vals = shard.gather_frames(m, v, dst=0)
    if rank == 0:
        q.put((rank, [x.tolist() for x in maps], [x.tolist() for x in vals]))
    else:
        q.put((rank, maps, vals))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_of_frames():
    """Config E's exchange step on CPU (gloo): rank 0 receives every rank's dense map and (ragged) vertex
    idepths; the other ranks receive nothing."""
    import math
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, maps, vals), (_, m1, v1) = res
    assert m1 is None and v1 is None
    assert len(maps) == 2 and math.isnan(maps[0][0][0]) and maps[0][1] == [1.0] * 4 and maps[1][2] == [2.0] * 4
    assert vals[0] == [0.0, 1.0, 2.0, 3.0, 4.0] and vals[1] == [100.0 + i for i in range(7)]


def test_single_process_is_identity():
    assert shard.max_over_ranks(1.5) == 1.5 and shard.sum_over_ranks(3) == 3.0
    assert shard.frames_for_rank(0, 1, 3) == [0, 1, 2]
    with pytest.raises(ValueError):
        shard.frames_for_rank(2, 2, 1)
