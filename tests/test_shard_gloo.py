"""World-size-2 checks of the sharding / statistics reduction on CPU (gloo)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from llicti_b200.shard import average_gradients, rank_share, reduce_stats, shard_range


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


# ---- data-parallel training: the gradient average and the loader's shares ---------------------------------------
def _toy(seed=0):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))


def _grad_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _toy()
        x = torch.arange(8 * 6, dtype=torch.float32).reshape(8, 6) / 10
        share = x[rank::world]                                  # this rank's half of the global batch
        model(share).mean().backward()                          # a mean over the batch, like the rate loss
        n = average_gradients(list(model.parameters()))
        if rank == 0:
            out.put((n, [p.grad.tolist() for p in model.parameters()]))       # (plain lists: the worker exits before the parent reads)
    finally:
        torch.distributed.destroy_process_group()


def test_two_rank_gradient_average_equals_the_global_batch():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    n, grads = out.get()
    model = _toy()
    x = torch.arange(8 * 6, dtype=torch.float32).reshape(8, 6) / 10
    model(x).mean().backward()
    assert n == sum(p.numel() for p in model.parameters())
    for g, p in zip(grads, model.parameters()):
        assert torch.allclose(torch.tensor(g), p.grad, atol=1e-6)


def test_gradient_average_is_identity_without_a_group():
    model = _toy()
    model(torch.ones(2, 6)).sum().backward()
    before = [p.grad.clone() for p in model.parameters()]
    assert average_gradients(list(model.parameters())) == 0
    assert all(torch.equal(a, p.grad) for a, p in zip(before, model.parameters()))


@pytest.mark.parametrize("n,world", [(12, 2), (13, 2), (100, 8), (3, 8), (16, 1)])
def test_rank_shares_are_disjoint_and_equal(n, world):
    order = list(torch.randperm(n, generator=torch.Generator().manual_seed(n)).tolist())
    shares = [rank_share(order, r, world) for r in range(world)]
    if n < world:
        assert all(s == order for s in shares)
        return
    assert len({len(s) for s in shares}) == 1 and len(shares[0]) == n // world
    flat = [i for s in shares for i in s]
    assert len(set(flat)) == len(flat) and set(flat) <= set(order)


def test_train_loader_shares_an_epoch_between_ranks(tmp_path):
    from PIL import Image
    import numpy as np
    from llicti_b200.image_dl import TrainImageLoader
    for i in range(6):
        Image.fromarray(np.full((40, 40, 3), 10 * i, np.uint8)).save(tmp_path / f"t{i}.png")     # image i is the constant 10 i
    seen = []
    for rank in range(2):
        dl = TrainImageLoader(str(tmp_path), 32, 2, seed=5, rank=rank, world=2)
        assert len(dl) == 2
        vals = [int(round(float(b[k, 0, 0, 0]) * 255)) // 10 for b in dl for k in range(b.shape[0])]
        assert len(vals) == 3
        seen += vals
    assert sorted(seen) == list(range(6))
