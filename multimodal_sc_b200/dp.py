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


def loss_scales(local_batch: int, world: int, c_sem: float, global_batch: int | None = None):
    """(inv_m for the per-sample PPO terms, gradient coefficient of the rank-local mean L_sem, this rank's weight in the
    global-minibatch mean L_sem).  Equal shards by default (global = local * world); ragged slices pass global_batch."""
    gb = local_batch * world if global_batch is None else int(global_batch)
    share = local_batch / gb if gb > 0 else 0.0
    return 1.0 / max(gb, 1), c_sem * share, share


def env_shard(n_local_envs: int, pg=None):
    """(first global environment index of this rank, global environment count): ranks own contiguous runs of whole
    environments in rank order; one small all-gather of the local counts per update."""
    world, rank = world_info(pg)
    if world == 1:
        return 0, n_local_envs
    counts = [None] * world
    dist.all_gather_object(counts, int(n_local_envs), group=pg)
    return sum(counts[:rank]), sum(counts)


def local_minibatch_indices(global_idx: torch.Tensor, n_envs_global: int, env_lo: int, n_envs_local: int):
    """The rank-local part of one GLOBAL minibatch ("sliced by rank so the union equals the 1-rank
    permutation").  global_idx: int64 slice of the global permutation over the flattened [T, N_global] rollout
    (flat index = t * N_global + env).  This rank owns environments [env_lo, env_lo + n_envs_local).
    Returns (rows, pos): rows = local flat indices (t * N_local + env - env_lo) of the owned samples, pos = their
    positions inside the global minibatch (the global sample index of the channel's noise counters).  Order follows
    the permutation, so one rank with every environment gets rows == global_idx, pos == arange."""
    env = global_idx % n_envs_global
    mine = (env >= env_lo) & (env < env_lo + n_envs_local)
    pos = torch.nonzero(mine).flatten()
    g = global_idx[pos]
    rows = (g // n_envs_global) * n_envs_local + (g % n_envs_global - env_lo)
    return rows.contiguous(), pos.contiguous()


def allreduce_flat_grads(g32: torch.Tensor, pg=None) -> torch.Tensor:
    world, _ = world_info(pg)
    if world > 1:
        dist.all_reduce(g32, op=dist.ReduceOp.SUM, group=pg)
    return g32


def grad_buckets(offsets, numel: int):
    """Contiguous regions of the flat gradient bucket in the order backward COMPLETES them.

    `offsets`: name -> (offset, n) in flat order (cam_enc.*, bev_enc.*, dec.*, ac.*).  Backward runs
    heads -> decoder -> channel -> camera encoder -> BEV encoder, so the [dec.*, ac.*] tail is final first, then the
    camera-encoder region, then the BEV-encoder region.  Returns [(name, lo, hi)] covering [0, numel) exactly once."""
    first = {}
    for name, (off, _) in offsets.items():
        top = name.split(".")[0]
        grp = top if top in ("cam_enc", "bev_enc") else "dec"
        first.setdefault(grp, off)
    cam, bev, dec = first["cam_enc"], first["bev_enc"], first["dec"]
    assert cam == 0 and cam < bev < dec < numel, (cam, bev, dec, numel)
    return [("dec", dec, numel), ("cam_enc", cam, bev), ("bev_enc", bev, dec)]


class OverlappedGradAllReduce:
    """All-reduce of the flat gradient bucket in the regions of grad_buckets(), each launched (async) as soon as
    backward has finished writing it, so the decoder's 2.5 M gradients cross NVLink while the encoders' backward
    kernels run.  ready(name) is called by Engine.backward on the compute stream; finish() joins before Adam.
    With world_size 1 every call is a no-op."""

    def __init__(self, g32: torch.Tensor, offsets, pg=None):
        self.g32, self.pg = g32, pg
        self.world, _ = world_info(pg)
        self.buckets = {name: (lo, hi) for name, lo, hi in grad_buckets(offsets, g32.numel())}
        self.pending = []
        self.done = set()

    def ready(self, name: str):
        if self.world == 1:
            return
        lo, hi = self.buckets[name]
        assert name not in self.done, f"bucket {name} reduced twice"
        self.done.add(name)
        self.pending.append(dist.all_reduce(self.g32[lo:hi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def finish(self):
        if self.world == 1:
            return
        assert self.done == set(self.buckets), f"buckets not reduced: {set(self.buckets) - self.done}"
        for w in self.pending:
            w.wait()                      # the current stream waits for the collective (no host block on NCCL)
        self.pending.clear()
        self.done.clear()


def global_adv_moments(sums: torch.Tensor, local_count: int, pg=None):
    """sums = float64 [sum A, sum A^2] over the local shard -> (global sums, global count).  The count is reduced with the
    sums (ranks may own different numbers of environments); one small all-reduce and one host read per rollout."""
    world, _ = world_info(pg)
    if world == 1:
        return sums, local_count
    buf = torch.cat([sums.reshape(2), torch.tensor([float(local_count)], dtype=sums.dtype, device=sums.device)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=pg)
    sums.copy_(buf[:2])
    return sums, int(round(float(buf[2])))
