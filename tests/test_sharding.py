"""CPU tests of the multi-GPU harness logic: the stream partition and the gloo world_size-2
reduction the bench uses for max-over-ranks timing (no collective touches frame data)."""
import os
import socket

import pytest

import __graft_entry__ as graft

pkg = graft.load_package()
sh = pkg.sharding


def test_stream_partition_is_a_partition():
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            mine = sh.shard_streams(256, world, r)
            assert all(s % world == r for s in mine)
            assert len(mine) == 256 // world
            seen += mine
        assert sorted(seen) == list(range(256))
    with pytest.raises(ValueError):
        sh.shard_streams(4, 2, 2)


def test_aggregate_fps_uses_slowest_rank():
    assert sh.aggregate_fps([(320, 10.0), (320, 20.0)]) == pytest.approx(640 / 0.020)
    assert sh.reduce_max(3.5) == 3.5
    assert sh.gather_records((1, 2, 3)) == [[1.0, 2.0, 3.0]]


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
    graft.load_package()
    mine = sh.shard_streams(256, world, rank)
    ms = 10.0 * (rank + 1)                       # rank 1 is the slow one
    dist.barrier()
    worst = sh.reduce_max(ms, dist)
    recs = sh.gather_records((len(mine) * 4, ms, rank), dist)
    q.put((rank, len(mine), worst, recs, sh.aggregate_fps(recs)))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world_size_2_reduction():
    import torch.multiprocessing as mp
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    port = _free_port()
    procs = [ctxmp.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, n_mine, worst, recs, fps in got:
        assert n_mine == 128
        assert worst == 20.0                      # MAX over ranks
        assert recs == [[512.0, 10.0, 0.0], [512.0, 20.0, 1.0]]
        assert fps == pytest.approx(1024 / 0.020)
