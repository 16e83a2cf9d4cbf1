"""N > 1 host-side logic on CPU: sample sharding and the all-reduce + rbo_finalize_sums merge, with two gloo ranks."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard(M, rank, world):
    b = (M * rank) // world
    return b, (M * (rank + 1)) // world - b


def _worker(rank, world, port, M, d, nth, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    lib = g.load_package()._lib.load()
    rng = np.random.default_rng(11)  # every rank draws the same global per-trajectory table, then keeps its shard
    table = rng.standard_normal((1 + d + nth, M)) * 3.0 + 5.0
    b, n = shard(M, rank, world)
    mine = table[:, b:b + n]
    sums = np.zeros(1 + 3 * (1 + d + nth) + 2)  # rows, then [n_failed, watchdog] (rbo_partial_sums_device)
    sums[0] = n
    for r in range(1 + d + nth):  # what rbo_stats_kernel writes per handle: [n*mean, M2, n*mean^2]
        mu = mine[r].mean()
        sums[1 + 3 * r] = n * mu
        sums[2 + 3 * r] = ((mine[r] - mu) ** 2).sum()
        sums[3 + 3 * r] = n * mu * mu
    t = torch.from_numpy(sums)
    dist.all_reduce(t)  # the only collective of the path
    m, s = C.c_double(), C.c_double()
    gm, gs, tm, ts = np.zeros(d), np.zeros(d), np.zeros(nth), np.zeros(nth)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.rbo_finalize_sums(p(sums), d, nth, C.byref(m), C.byref(s), p(gm), p(gs), p(tm), p(ts)) == 0
    ok = (np.isclose(m.value, table[0].mean(), rtol=1e-13) and np.isclose(s.value, table[0].std(ddof=1), rtol=1e-11)
          and np.allclose(gm, table[1:1 + d].mean(axis=1), rtol=1e-13) and np.allclose(gs, table[1:1 + d].std(axis=1, ddof=1), rtol=1e-11)
          and np.allclose(ts, table[1 + d:].std(axis=1, ddof=1), rtol=1e-11))
    ret[rank] = (bool(ok), b, n)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_merge(pkg):  # the `pkg` fixture builds librbo.so on a clean checkout before the workers load it
    world, M, d, nth = 2, 1001, 5, 1
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, M, d, nth, ret), nprocs=world, join=True)
    assert all(ret[r][0] for r in range(world))
    # shards are contiguous, disjoint and cover every sample exactly once
    spans = sorted((ret[r][1], ret[r][2]) for r in range(world))
    assert spans[0][0] == 0 and spans[0][0] + spans[0][1] == spans[1][0] and spans[1][0] + spans[1][1] == M


def test_shard_covers_all_samples():
    for M in (1, 7, 16384, 65536):
        for world in (1, 2, 4, 8):
            tot, prev = 0, 0
            for r in range(world):
                b, n = shard(M, r, world)
                assert b == prev
                prev = b + n
                tot += n
            assert tot == M
