"""Multi-rank host logic on CPU: world_size 2 over gloo (127.0.0.1)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fec import pkg

import importlib
dispatch = importlib.import_module("sdrpp-dvbs-demodulator_b200.dispatch")


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 16, 4096, 4097):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                f0, f1 = dispatch.shard_range(n, world, r)
                assert 0 <= f0 <= f1 <= n
                seen.extend(range(f0, f1))
            assert seen == list(range(n))


def test_merge_in_order():
    shares = [(4, ["e", "f"]), (0, ["a", "b", "c", "d"])]
    assert dispatch.merge_in_order(shares) == list("abcdef")


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f0, f1 = dispatch.shard_range(101, world, rank)
    ms = dispatch.reduce_max([10.0 + rank, 3.0 - rank])
    tot = dispatch.reduce_sum([f1 - f0])
    q.put((rank, f0, f1, ms, tot))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reduce_over_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1:3] == (0, 51) and got[1][1:3] == (51, 101)
    for g in got:
        assert g[3] == [11.0, 3.0]      # max over ranks, what bench.py reports as the step time
        assert g[4] == [101.0]
