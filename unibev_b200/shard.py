"""Batch sharding for multi-GPU runs: nuScenes samples are independent through the whole hot path
(SURVEY.md 8e), so each rank takes a contiguous slice of the sample list and no data-path collective
exists.  The only cross-rank traffic is bookkeeping (barrier, max-over-ranks of the timed duration)."""
import torch
import torch.distributed as dist


def shard_range(n_samples, rank, world):
    """Contiguous, balanced split: first (n % world) ranks get one extra sample."""
    if not (0 <= rank < world):
        raise ValueError(f'rank {rank} outside world of {world}')
    base, extra = divmod(n_samples, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def max_over_ranks(value, device='cpu'):
    """Largest `value` over all ranks (identity when not distributed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device='cpu'):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
