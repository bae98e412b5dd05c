"""World-size-2 checks of the sharding / statistics reduction on CPU (gloo)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from llicti_b200.shard import reduce_stats, shard_range


@pytest.mark.parametrize("n,world", [(24, 1), (24, 2), (25, 2), (100, 8), (3, 8), (0, 4), (512, 8)])
def test_shards_partition_the_images(n, world):
    seen = []
    for r in range(world):
        a, b = shard_range(n, r, world)
        assert 0 <= a <= b <= n
        seen.extend(range(a, b))
    assert seen == list(range(n))
    sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
    assert max(sizes) == -(-n // world)
    with pytest.raises(ValueError):
        shard_range(n, world, world)


def test_reduce_is_identity_without_a_group():
    s, m = reduce_stats([1.0, 2.0], [3.0])
    assert s == [1.0, 2.0] and m == [3.0]


def _worker(rank, world, port, n_images, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = shard_range(n_images, rank, world)
        pixels = float((b - a) * 512 * 768)
        nbytes = float(sum(1000 + i for i in range(a, b)))      # stand-in for the compressed size of image i
        seconds = 0.5 + 0.25 * rank                              # per-rank device time: the job takes the max
        sums, maxs = reduce_stats([pixels, nbytes], [seconds])
        if rank == 0:
            out.put((sums, maxs))
    finally:
        torch.distributed.destroy_process_group()


def test_two_rank_reduction_over_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    n_images, world = 25, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    sums, maxs = out.get()
    assert sums[0] == n_images * 512 * 768
    assert sums[1] == sum(1000 + i for i in range(n_images))
    assert maxs == [0.75]
