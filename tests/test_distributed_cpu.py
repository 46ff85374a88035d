"""CPU, world_size 2 over gloo: the sharding / final-gather plumbing of the N>1 path.  The render itself is
a stand-in here (the CUDA path cannot run on the CPU box); on GPUs the same code runs with NCCL."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quantumdistortion_b200.distributed import render_sharded, shard_bounds


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 2, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _worker(rank, world, port, n_clips, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = torch.arange(n_clips * 5, dtype=torch.float32).reshape(n_clips, 5)
        seen = []

        def fake_render(shard):  # what a rank would hand to the GPU
            seen.append(tuple(shard.shape))
            return shard * 2.0 + 1.0

        y, (lo, hi) = render_sharded(x, fake_render, gather_to=0)
        ok = seen == [(hi - lo, 5)]
        if rank == 0:
            ok = ok and torch.equal(y, x * 2.0 + 1.0)
        else:
            ok = ok and torch.equal(y, x[lo:hi] * 2.0 + 1.0)
        y2, _ = render_sharded(x, fake_render, gather_to=None)
        ok = ok and torch.equal(y2, x[lo:hi] * 2.0 + 1.0)
        q.put((rank, bool(ok), lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 7, 1])
def test_render_sharded_two_ranks_gloo(n_clips):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_clips
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == n_clips
