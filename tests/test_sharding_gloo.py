"""world_size-2 gloo test (CPU) of the multi-GPU host logic: pair-batch sharding with no data-path
collective, max-over-ranks timing, whole-job throughput aggregation."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kangaroo_b200.sharding import aggregate_throughput, reduce_max, shard_pairs


def test_shard_pairs_partitions_every_batch():
    for n in (0, 1, 7, 8, 512, 513):
        for world in (1, 2, 3, 4, 8):
            got = [i for r in range(world) for i in shard_pairs(n, world, r)]
            assert got == list(range(n))
            sizes = [len(shard_pairs(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_pairs(n_pairs, world, rank)
    # every rank "processes" its shard at a rank-dependent speed; the job time is the slowest rank's
    seconds = 1.0 + rank
    gathered = [None] * world
    dist.all_gather_object(gathered, list(mine))
    tmax = reduce_max(seconds)
    thr = aggregate_throughput(len(mine), seconds)
    dist.barrier()
    if rank == 0:
        q.put((gathered, tmax, thr))
    dist.destroy_process_group()


def test_two_process_sharding_and_timing_reduction():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_pairs, world = 11, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, tmax, thr = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(i for g in gathered for i in g) == list(range(n_pairs))  # disjoint and complete
    assert tmax == 2.0                                                      # max over ranks, not mean
    assert thr == pytest.approx(n_pairs / 2.0)
