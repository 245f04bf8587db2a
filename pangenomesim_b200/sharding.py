"""Gene sharding across engines / ranks.

Columns of the alignment evolve independently given the tree, and every random draw is keyed by
(seed, edge, gene_id, block), so a rank that owns a contiguous range of the gene table computes
exactly the rows a single engine would.  The sweep needs no collective; only the partial
mismatch matrices of kernel K2 are summed across ranks (all-reduce, uint32 counts)."""
from __future__ import annotations


def shard_range(n_genes: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [begin, end) range of the gene table owned by `rank`."""
    if not 0 <= rank < world:
        raise ValueError("rank outside [0, world)")
    base, extra = divmod(n_genes, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def aggregate_throughput(units_per_rank, seconds_per_rank) -> float:
    """Whole-job throughput: all units over the slowest rank's time."""
    return float(sum(units_per_rank)) / max(seconds_per_rank)


def allreduce_mismatch(partial, group=None):
    """Sum per-rank partial n x n mismatch counts in place (torch tensor on the rank's device;
    NCCL on GPUs, gloo in the CPU tests).  uint32 counts are carried as int32/int64 tensors."""
    import torch.distributed as dist
    dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


def mismatch_tile_words(n_genomes: int) -> int:
    """uint32 words of the packed upper-triangular 64 x 64 tiles K2 keeps on the device (include/pgs.h)"""
    t = (n_genomes + 63) // 64
    return t * (t + 1) // 2 * 4096


def tiles_to_square(tiles, n_genomes: int):
    """Packed upper-triangular tiles (numpy uint32, e.g. after an all-reduce of the engines' buffers) ->
    the symmetric n x n matrix with a zero diagonal that pgs_mismatch returns."""
    import numpy as np
    t = (n_genomes + 63) // 64
    tiles = np.asarray(tiles, dtype=np.uint32).reshape(t * (t + 1) // 2, 64, 64)
    full = np.zeros((t * 64, t * 64), dtype=np.uint32)
    k = 0
    for ti in range(t):
        for tj in range(ti, t):
            block = tiles[k] if ti != tj else np.triu(tiles[k], 1)
            full[64 * ti:64 * ti + 64, 64 * tj:64 * tj + 64] = block
            k += 1
    full = full[:n_genomes, :n_genomes]
    return full + full.T
