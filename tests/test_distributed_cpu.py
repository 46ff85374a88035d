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


# ------------------------------------------------------------------ random spectral FX under sharding
_FX_KW = dict(use_multiband=True, spectral_fx_mode="bin_scramble", spectral_fx_strength=0.55)


def _fx_resolved(n=3000):
    from quantumdistortion_b200.pipeline import _resolve_kwargs
    res, _ = _resolve_kwargs(n, 48000, 2048, dict(_FX_KW))
    assert res.fx_rng == ("pick", 4)
    return res


def test_fx_tables_sharded_equal_unsharded():
    """Every clip of a shard gets the np.random draws it would get in the unsharded render (seeds=None), a per-clip
    seed list is sliced per shard, and the global state ends where the unsharded render leaves it."""
    import numpy as np
    from quantumdistortion_b200.pipeline import fx_tables
    res = _fx_resolved()
    np.random.seed(9)
    full = fx_tables(res, 5, None)
    end_state = np.random.get_state()[1].copy()
    for lo, hi in ((0, 3), (3, 5), (2, 2)):
        np.random.seed(9)
        part = fx_tables(res, hi - lo, None, shard=(lo, hi, 5))
        assert len(part) == hi - lo and all(np.array_equal(a, b) for a, b in zip(part, full[lo:hi]))
        assert np.array_equal(np.random.get_state()[1], end_state)
    seeded = fx_tables(res, 5, [11, 12, 13, 14, 15])
    part = fx_tables(res, 2, [11, 12, 13, 14, 15], shard=(3, 5, 5))
    assert all(np.array_equal(a, b) for a, b in zip(part, seeded[3:5]))
    assert np.array_equal(fx_tables(res, 2, [14, 15], shard=(3, 5, 5))[1], seeded[4])   # already sliced by the caller
    with pytest.raises(ValueError):
        fx_tables(res, 2, [1, 2, 3], shard=(3, 5, 5))
    assert len(fx_tables(res, 4, 77, shard=(0, 4, 8))) == 1                             # one shared table


def _fx_worker(rank, world, port, q):
    import numpy as np
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from quantumdistortion_b200 import distributed, pipeline
        res = _fx_resolved()

        def fake_process_batch(part, sr, seeds=None, shard=None, **kw):   # stands in for the GPU: returns what it drew
            tabs = pipeline.fx_tables(res, int(part.shape[0]), seeds, shard)
            rows = [torch.from_numpy(t[0, 0, :16].astype(np.float32)) for t in tabs]
            return (torch.stack(rows) if rows else torch.zeros((0, 16))), None

        pipeline.process_batch = fake_process_batch
        x = torch.zeros((5, 3000))
        out = {}
        for tag, seeds in (("none", None), ("list", [21, 22, 23, 24, 25])):
            np.random.seed(1234)
            y, _ = distributed.process_batch_sharded(x, 48000, seeds=seeds, **_FX_KW)
            np.random.seed(1234)
            ref, _ = fake_process_batch(x, 48000, seeds=seeds)
            out[tag] = bool(torch.equal(y, ref)) if rank == 0 else True
        q.put((rank, out["none"], out["list"]))
    finally:
        dist.destroy_process_group()


def test_process_batch_sharded_fx_seeds_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_fx_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, True), (1, True, True)]
