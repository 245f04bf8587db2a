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
