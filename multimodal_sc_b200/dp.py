"""Host-side data-parallel logic, device-agnostic so it is testable with gloo on CPU.

Pure data parallelism: ranks own whole trajectories (the GAE scan stays rank-local); the only
exchanges are (1) one all-reduce(sum) of the flat fp32 gradient bucket per optimizer step and
(2) one all-reduce of (sum A, sum A^2) per rollout for the global advantage normalisation.
Per-sample loss terms are pre-scaled by 1/global_batch and L_sem by 1/world so that the SUM of
the ranks' gradients is the gradient of the global-minibatch mean; channel noise counters use
the global sample index (`sample_offset`) so the N-rank step equals the 1-rank step."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_info(pg=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(pg), dist.get_rank(pg)
    return 1, 0


def shard_bounds(n: int, world: int, rank: int):
    """contiguous, balanced [lo, hi) split of n items (whole environments / minibatch rows)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def loss_scales(local_batch: int, world: int, c_sem: float):
    """(inv_m for per-sample PPO terms, coefficient for the rank-local mean L_sem)."""
    return 1.0 / (local_batch * world), c_sem / world


def allreduce_flat_grads(g32: torch.Tensor, pg=None) -> torch.Tensor:
    world, _ = world_info(pg)
    if world > 1:
        dist.all_reduce(g32, op=dist.ReduceOp.SUM, group=pg)
    return g32


def global_adv_moments(sums: torch.Tensor, local_count: int, pg=None):
    """sums = float64 [sum A, sum A^2] over the local shard -> (global sums, global count)."""
    world, _ = world_info(pg)
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=pg)
    return sums, local_count * world
