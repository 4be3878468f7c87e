"""Host logic of the image-parallel path (N > 1), on CPU: shard arithmetic and the
one-process-per-rank front end with a world-size-2 gloo group (no GPU, no libudal compute)."""
import os
import socket

import numpy as np
import pytest


def test_shard_ranges_cover_the_batch_in_order():
    import udal_b200 as u
    sr = u.scheduler.shard_ranges
    assert sr(64, 1) == [(0, 64)]
    assert sr(64, 8) == [(i * 8, i * 8 + 8) for i in range(8)]
    assert sr(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert sr(3, 8) == [(0, 1), (1, 2), (2, 3)] + [(3, 3)] * 5
    assert sr(0, 2) == [(0, 0), (0, 0)]
    for b in range(0, 40):
        for w in range(1, 9):
            r = sr(b, w)
            assert len(r) == w and r[0][0] == 0 and r[-1][1] == b
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(e - s for s, e in r) == -(-b // w) if b else True
    with pytest.raises(ValueError):
        sr(4, 0)


def test_concat_results_skips_empty_shards():
    import udal_b200 as u
    a = (np.arange(6).reshape(3, 2), np.arange(3))
    b = (np.arange(4).reshape(2, 2) + 100, np.arange(2) + 100)
    out = u.scheduler.concat_results([a, None, b])
    assert out[0].shape == (5, 2) and out[1].tolist() == [0, 1, 2, 100, 101]
    assert u.scheduler.concat_results([None, None]) is None


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, batch, q):
    import torch.distributed as dist
    import udal_b200 as u
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)  # every rank sees the same host batch
    levels = [rng.normal(size=(batch, 4 >> l or 1, 4 >> l or 1, 3)).astype(np.float32) for l in range(3)]
    scales = np.arange(batch, dtype=np.float32) + 1

    def fn(shard, extra):  # stands in for HeadSampler.detect on this rank's GPU
        s = extra[0]
        return (np.stack([x.reshape(x.shape[0], -1).sum(1) for x in shard], 1) * s[:, None], s.copy())

    out = u.scheduler.run_sharded(fn, levels, batch, rank, world, extra=[scales], gather=True)
    mine = u.scheduler.run_sharded(fn, levels, batch, rank, world, extra=[scales], gather=False)
    ref = fn(levels, [scales])
    ok = True
    if rank == 0:
        ok = np.array_equal(out[0], ref[0]) and np.array_equal(out[1], ref[1])
    else:
        ok = out is None
    s, e = u.scheduler.shard_ranges(batch, world)[rank]
    if e > s:
        ok = ok and np.array_equal(mine[1], scales[s:e])
    else:
        ok = ok and mine is None
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [7, 1])
def test_run_sharded_world2_gloo(batch):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
