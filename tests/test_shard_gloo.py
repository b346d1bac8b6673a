"""world_size-2 gloo test (CPU) of the multi-GPU plan: batch shards tile the batch exactly, and the
job time / throughput reduction is the max over ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from shard import aggregate_tflops, reduce_max_time, shard_batch


@pytest.mark.parametrize("total,world", [(64, 1), (64, 2), (64, 4), (64, 8), (5, 2), (3, 8), (0, 2)])
def test_shards_tile_the_batch(total, world):
    covered = []
    for r in range(world):
        s, c = shard_batch(total, world, r)
        covered.extend(range(s, s + c))
    assert covered == list(range(total))
    counts = [shard_batch(total, world, r)[1] for r in range(world)]
    assert max(counts) - min(counts) <= 1


def test_bad_requests_raise():
    with pytest.raises(ValueError):
        shard_batch(8, 0, 0)
    with pytest.raises(ValueError):
        shard_batch(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total_b = 5
        start, count = shard_batch(total_b, world, rank)
        # each rank "runs" its shard: result rows tagged by global batch index
        local = torch.arange(start, start + count, dtype=torch.float64)
        gathered = [None] * world
        dist.all_gather_object(gathered, local.tolist())
        local_ms = 10.0 * (rank + 1)  # rank 1 is slower
        job_ms = reduce_max_time(local_ms, dist)
        tf = aggregate_tflops(2e12, local_ms, dist)
        out[rank] = (gathered, job_ms, tf)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_and_max_time():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        gathered, job_ms, tf = out[r]
        assert sorted(sum(gathered, [])) == [0.0, 1.0, 2.0, 3.0, 4.0]
        assert job_ms == 20.0                      # max over ranks, not the local time
        assert tf == pytest.approx(2e12 / 20e-3 / 1e12)


def test_single_process_reduce_is_identity():
    assert reduce_max_time(3.5) == 3.5
    assert aggregate_tflops(1e12, 1000.0) == pytest.approx(1.0)
