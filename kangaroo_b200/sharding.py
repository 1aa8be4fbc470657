"""Multi-GPU host logic: stereo pairs are independent, so a batch is split across one process per GPU
with NO data-path collective.  torch.distributed is used only for the timing barrier and
the max-over-ranks reduction of the measured time."""
from __future__ import annotations


def shard_pairs(n_pairs: int, world: int, rank: int) -> range:
    """Contiguous block of pair indices for `rank` (blocks differ by at most one pair)."""
    base, rem = divmod(n_pairs, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def reduce_max(value: float, device=None) -> float:
    """Max over ranks of a host float (device = where the backend wants the tensor: cuda for nccl)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(pairs_this_rank: int, seconds_this_rank: float, device=None) -> float:
    """Whole-job pairs/s: all pairs of all ranks divided by the slowest rank's time."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return pairs_this_rank / seconds_this_rank
    n = torch.tensor([float(pairs_this_rank)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(n.item()) / reduce_max(seconds_this_rank, device)
