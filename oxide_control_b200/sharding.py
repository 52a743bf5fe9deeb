"""Host-side multi-GPU logic: environments shard across ranks with no data-path collective;
`torch.distributed` is used only to gather a small per-rank statistics vector at report boundaries."""
from __future__ import annotations

from typing import Dict, Sequence, Tuple


def shard_range(rank: int, world: int, envs_per_rank: int) -> Tuple[int, int]:
    """Global env ids [lo, hi) owned by `rank`. Global ids key the Philox control stream, so a trajectory does not
    depend on how many GPUs the job runs on."""
    if not (0 <= rank < world) or envs_per_rank < 1:
        raise ValueError("bad shard arguments")
    return rank * envs_per_rank, (rank + 1) * envs_per_rank


def owner_of(global_env: int, envs_per_rank: int) -> Tuple[int, int]:
    """(rank, local index) of a global env id."""
    return global_env // envs_per_rank, global_env % envs_per_rank


STAT_KEYS = ("env_steps", "sum_ncon", "sum_nefc", "sum_niter", "diverged", "episode_return_sum", "episodes")


def gather_stats(local: Dict[str, float], device=None) -> Dict[str, float]:
    """Sum the per-rank statistics vector over all ranks (all_reduce of len(STAT_KEYS) doubles; NCCL on GPUs,
    gloo in the CPU tests). Returns the global totals on every rank. Works without an initialised process group."""
    import torch
    import torch.distributed as dist
    vec = torch.tensor([float(local.get(k, 0.0)) for k in STAT_KEYS], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return {k: float(v) for k, v in zip(STAT_KEYS, vec.tolist())}


def max_over_ranks(x: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
