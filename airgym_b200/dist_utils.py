"""Env-axis sharding helpers: envs are independent, so ranks own contiguous blocks and the only
cross-rank traffic on the env path is the timing reduction of the bench."""
import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard(num_envs_total: int, rank: int, world: int):
    """Contiguous block [offset, offset+count) of a global env axis; the first `rem` ranks take one extra env."""
    base, rem = divmod(num_envs_total, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def max_over_ranks(value: float, dist=None, device="cpu") -> float:
    """Time-like quantities are reported as the max over ranks (never a wall clock on rank 0)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
