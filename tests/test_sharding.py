"""CPU tests of the multi-rank host logic (world_size 2, gloo): gene shard ranges, whole-job
aggregation and the all-reduce of partial mismatch matrices (the only collective of the path)."""
import os
import socket

import numpy as np
import pytest

from pangenomesim_b200.sharding import aggregate_throughput, shard_range


def test_shard_ranges_partition_the_table():
    for G in (0, 1, 7, 5000, 5001):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(G, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == G
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_aggregate_is_units_over_slowest_rank():
    assert aggregate_throughput([10, 10], [1.0, 2.0]) == 10.0


def _worker(rank, world, port, G, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_py
    from pangenomesim_b200.sharding import allreduce_mismatch, shard_range
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import treegen
    n, L, seed = 6, 100, 3
    parent, rate, leaf = treegen.coalescent(n, seed=1, mut_rate=0.3)
    root = len(parent) - 1
    g0, g1 = shard_range(G, rank, world)
    part = torch.zeros((n, n), dtype=torch.int64)
    for g in range(g0, g1):
        rows = oracle_py.gene_all_nodes(seed, parent, rate, root, g, L)
        for i in range(n):
            for j in range(n):
                part[i, j] += oracle_py.mismatch(rows[i], rows[j])
    allreduce_mismatch(part)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        q.put((part.numpy(), float(t.item())))
    dist.destroy_process_group()


def test_two_rank_mismatch_allreduce_equals_single_rank():
    import torch.multiprocessing as mp
    from oracle import oracle_py
    import treegen
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    G = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, G, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, L, seed = 6, 100, 3
    parent, rate, leaf = treegen.coalescent(n, seed=1, mut_rate=0.3)
    want = np.zeros((n, n), dtype=np.int64)
    for g in range(G):
        rows = oracle_py.gene_all_nodes(seed, parent, rate, len(parent) - 1, g, L)
        for i in range(n):
            for j in range(n):
                want[i, j] += oracle_py.mismatch(rows[i], rows[j])
    assert np.array_equal(got, want) and tmax == 2.0


def test_packed_tiles_round_trip():
    """the packed upper-triangular layout of K2's counts (include/pgs.h) and its unpacking"""
    import numpy as np
    from pangenomesim_b200.sharding import mismatch_tile_words, tiles_to_square
    for n in (1, 2, 63, 64, 65, 130, 200):
        rng = np.random.default_rng(n)
        up = np.triu(rng.integers(0, 1000, size=(n, n)).astype(np.uint32), 1)
        want = up + up.T
        t = (n + 63) // 64
        tiles = np.zeros((t * (t + 1) // 2, 64, 64), np.uint32)
        k = 0
        for ti in range(t):
            for tj in range(ti, t):
                blk = up[64 * ti:64 * ti + 64, 64 * tj:64 * tj + 64]
                tiles[k, :blk.shape[0], :blk.shape[1]] = blk
                k += 1
        assert tiles.size == mismatch_tile_words(n)
        assert np.array_equal(tiles_to_square(tiles.reshape(-1), n), want)
        # partial results of two engines add up tile by tile
        assert np.array_equal(tiles_to_square(tiles * 2, n), tiles_to_square(tiles, n) * 2)
