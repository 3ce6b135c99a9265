"""world_size-2 gloo test of the data-parallel host logic (multimodal_sc_b200/dp.py) on CPU.

Invariant: with per-sample terms scaled by 1/global_batch, L_sem by 1/world,
noise counters on the global sample index and advantage moments all-reduced, the SUM over ranks
of the shard gradients equals the 1-rank gradient on the concatenated minibatch.  The arithmetic
is done by the oracle here (CPU); the native agent uses the same dp.py helpers around its kernels."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodal_sc_b200 import dp, synthetic
from multimodal_sc_b200.spec import SpecV0

SPEC = SpecV0()


def _grads(policy, mb, lo, hi, inv_m, sem_coef, seed):
    from oracle.ppo import ppo_loss
    import dataclasses
    policy.zero_grad(set_to_none=True)
    sl = slice(lo, hi)
    logits, value, aux = policy(mb["cam"][sl], mb["bev"][sl], 10.0, channel="rayleigh", seed=seed, sample_offset=lo)
    sp1 = dataclasses.replace(SPEC, c_sem=0.0)
    out = ppo_loss(logits, mb["actions"][sl], mb["logp_old"][sl], mb["adv"][sl], value, mb["ret"][sl],
                   torch.tensor(0.0), sp1)
    n_local = hi - lo
    total = out["total"] * (n_local * inv_m) + sem_coef * aux["l_sem"]     # means -> sums / global batch
    total.backward()
    return torch.cat([p.grad.flatten() for p in policy.parameters()])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from oracle.model import SemComPolicy
        from oracle import ppo as oppo
        assert dp.world_info() == (world, rank)
        B = 6
        mb = synthetic.minibatch(SPEC, B, seed=0)
        pol = SemComPolicy(SPEC)
        pol.reset_parameters(std=0.05, seed=1)
        lo, hi = dp.shard_bounds(B, world, rank)
        inv_m, sem_coef, _ = dp.loss_scales(hi - lo, world, SPEC.c_sem)
        g = _grads(pol, mb, lo, hi, inv_m, sem_coef, seed=42)
        dp.allreduce_flat_grads(g)
        # advantage moments
        ro = synthetic.rollout(SPEC, 8, 4, seed=3)
        elo, ehi = dp.shard_bounds(4, world, rank)
        adv, _ = oppo.gae(ro["rewards"][:, elo:ehi], ro["values"][:, elo:ehi], ro["dones"][:, elo:ehi],
                          ro["last_value"][elo:ehi], SPEC.gamma, SPEC.lam)
        sums = torch.tensor([adv.double().sum(), adv.double().pow(2).sum()])
        sums, count = dp.global_adv_moments(sums, adv.numel())
        if rank == 0:
            q.put((g.numpy(), sums.numpy(), count))   # by value: the worker exits right after
    finally:
        dist.destroy_process_group()


def test_two_rank_dp_equals_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    g2, sums, count = q.get(timeout=600)
    g2, sums = torch.from_numpy(g2), torch.from_numpy(sums)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    from oracle.model import SemComPolicy
    from oracle import ppo as oppo
    torch.set_num_threads(4)
    mb = synthetic.minibatch(SPEC, 6, seed=0)
    pol = SemComPolicy(SPEC)
    pol.reset_parameters(std=0.05, seed=1)
    g1 = _grads(pol, mb, 0, 6, 1.0 / 6, SPEC.c_sem, seed=42)
    assert torch.allclose(g2, g1, rtol=2e-4, atol=2e-6), float((g2 - g1).abs().max())
    ro = synthetic.rollout(SPEC, 8, 4, seed=3)
    adv, _ = oppo.gae(ro["rewards"], ro["values"], ro["dones"], ro["last_value"], SPEC.gamma, SPEC.lam)
    assert count == adv.numel()
    assert torch.allclose(sums, torch.tensor([adv.double().sum(), adv.double().pow(2).sum()]), rtol=1e-9)


def test_shard_bounds_cover_everything():
    for n in (1, 7, 8, 4096):
        for w in (1, 2, 3, 8):
            spans = [dp.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


# ---- bucketed all-reduce launched as backward completes each region -------------------------
def test_grad_buckets_cover_flat_bucket_in_backward_order():
    from multimodal_sc_b200.model import FlatParams
    flat = FlatParams(SPEC, "cpu")
    b = dp.grad_buckets(flat.offsets, flat.numel)
    assert [n for n, _, _ in b] == ["dec", "cam_enc", "bev_enc"]          # the order backward finishes them
    spans = sorted((lo, hi) for _, lo, hi in b)
    assert spans[0][0] == 0 and spans[-1][1] == flat.numel
    assert all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))   # no gap, no overlap
    reg = {n: (lo, hi) for n, lo, hi in b}
    for name, (off, n) in flat.offsets.items():
        top = name.split(".")[0]
        lo, hi = reg[top if top in ("cam_enc", "bev_enc") else "dec"]
        assert lo <= off and off + n <= hi, name


def _overlap_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multimodal_sc_b200.model import FlatParams
        flat = FlatParams(SPEC, "cpu")
        g = torch.Generator().manual_seed(7 + rank)
        flat.g32.copy_(torch.randn(flat.numel, generator=g))
        whole = flat.g32.clone()
        ov = dp.OverlappedGradAllReduce(flat.g32, flat.offsets)
        for name in ("dec", "cam_enc", "bev_enc"):                         # Engine.backward's call order
            ov.ready(name)
        ov.finish()
        dp.allreduce_flat_grads(whole)
        ok = torch.equal(flat.g32, whole)
        # a second step reuses the object; a bucket announced twice or a missing bucket must be refused
        ov.ready("dec")
        try:
            ov.ready("dec")
            twice = False
        except AssertionError:
            twice = True
        try:
            ov.finish()
            missing = False
        except AssertionError:
            missing = True
        if rank == 0:
            q.put((ok, twice, missing))
    finally:
        dist.destroy_process_group()


def test_overlapped_bucket_allreduce_equals_single_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, twice, missing = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert ok and twice and missing


# ---- update()-level equivalence with the rank-sliced GLOBAL permutation ---------------------------
def _update_worker(rank, world, port, q):
    """each rank owns an UNEQUAL run of whole environments (1 and 3 of 4), draws the same global permutation as the
    1-rank oracle and processes, of every global minibatch, the samples it owns at their global noise positions —
    the host logic of multimodal_sc_b200.ppo.PPOAgent.update, with the oracle doing the arithmetic on the CPU."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        import dataclasses
        from oracle.model import SemComPolicy
        from oracle import ppo as oppo
        T, N, MB = 4, 4, 8
        ro = synthetic.rollout(SPEC, T, N, seed=5)
        env_bounds = [(0, 1), (1, 4)][rank]
        n_loc = env_bounds[1] - env_bounds[0]
        loc = {k: (v[:, env_bounds[0]:env_bounds[1]] if v.dim() >= 2 else v[env_bounds[0]:env_bounds[1]]) for k, v in ro.items()}
        env_lo, n_glob = dp.env_shard(n_loc)
        assert (env_lo, n_glob) == (env_bounds[0], N)
        pol = SemComPolicy(SPEC)
        pol.reset_parameters(std=0.05, seed=1)
        opt = torch.optim.Adam(pol.parameters(), lr=SPEC.lr, betas=(SPEC.beta1, SPEC.beta2), eps=SPEC.adam_eps)
        # advantages: rank-local scan, global normalisation moments
        adv, ret = oppo.gae(loc["rewards"], loc["values"], loc["dones"], loc["last_value"], SPEC.gamma, SPEC.lam)
        sums = torch.tensor([adv.double().sum(), adv.double().pow(2).sum()])
        sums, count = dp.global_adv_moments(sums, adv.numel())
        assert count == T * N                                  # unequal shards: the count is reduced, not local * world
        mean = sums[0] / count
        std = (sums[1] / count - mean * mean).clamp_min(0).sqrt()
        adv = ((adv.double() - mean) / (std + SPEC.adv_eps)).float()
        flat = lambda x: x.reshape(T * n_loc, *x.shape[2:])
        cam, bev, act, lp = flat(loc["cam"]), flat(loc["bev"]), flat(loc["actions"]), flat(loc["logp"])
        adv, ret = adv.reshape(-1), ret.reshape(-1)
        gen = torch.Generator().manual_seed(1234)
        sp0 = dataclasses.replace(SPEC, c_sem=0.0)
        step = 0
        covered = []
        perm = torch.randperm(T * N, generator=gen)
        for i in range(0, T * N - MB + 1, MB):
            rows, pos = dp.local_minibatch_indices(perm[i:i + MB], n_glob, env_lo, n_loc)
            covered.append(pos)
            inv_m, sem_coef, share = dp.loss_scales(rows.numel(), world, SPEC.c_sem, global_batch=MB)
            opt.zero_grad(set_to_none=True)
            logits, value, aux = pol(cam[rows], bev[rows], 10.0, channel="rayleigh", seed=1234 + step, sample_idx=pos)
            out = oppo.ppo_loss(logits, act[rows], lp[rows], adv[rows], value, ret[rows], torch.tensor(0.0), sp0)
            (out["total"] * (rows.numel() * inv_m) + sem_coef * aux["l_sem"]).backward()
            g = torch.cat([p.grad.flatten() for p in pol.parameters()])
            dp.allreduce_flat_grads(g)
            off = 0
            for p in pol.parameters():
                p.grad = g[off:off + p.numel()].view_as(p).clone()
                off += p.numel()
            opt.step()
            step += 1
        q.put((rank, perm.numpy(), [c.numpy() for c in covered],
               torch.cat([p.detach().flatten() for p in pol.parameters()]).numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_update_equals_single_rank_update_on_concatenated_rollout():
    from oracle.model import SemComPolicy
    from oracle.ppo import PPOAgent
    import numpy as np
    T, N, MB = 4, 4, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + 7
    procs = [ctx.Process(target=_update_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=600) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the 1-rank reference: oracle update() on the whole rollout
    pol = SemComPolicy(SPEC)
    pol.reset_parameters(std=0.05, seed=1)
    ag = PPOAgent(pol, SPEC, minibatch=MB, channel="rayleigh", seed=1234)
    ag.update(synthetic.rollout(SPEC, T, N, seed=5))
    ref = torch.cat([p.detach().flatten() for p in pol.parameters()]).numpy()
    # same permutation on both ranks and in the oracle (bit-exact), and the rank slices partition every minibatch
    assert np.array_equal(got[0][1], got[1][1]) and np.array_equal(got[0][1], ag.last_perms[0].numpy())
    for c0, c1 in zip(got[0][2], got[1][2]):
        assert sorted(np.concatenate([c0, c1]).tolist()) == list(range(MB))
    # both ranks hold the same parameters, equal to the 1-rank update (two Adam steps: differences are fp32 rounding
    # of differently ordered gradient sums, amplified by at most lr per step where a gradient is ~0)
    assert np.array_equal(got[0][3], got[1][3])
    diff = np.abs(got[0][3] - ref)
    assert np.mean(diff > 0.02 * SPEC.lr) < 5e-3, float(np.mean(diff > 0.02 * SPEC.lr))
    assert diff.max() <= 2.1 * 2 * SPEC.lr
