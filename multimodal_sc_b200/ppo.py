"""Native PPO update: GAE scan, fused loss, explicit backward, flat all-reduce, fused Adam.

Drop-in for `oracle/ppo.py: PPOAgent` (same constructor arguments, `update(rollout) -> stats`,
`minibatch_step(...)`; upstream /root/reference/README.md:1-2 has no
trainer).  One optimizer step is a fixed kernel schedule on the current stream:

    sc_step_begin (zero flat grad + stats, step counter) -> Engine.forward -> sc_mse_fwd_bwd -> sc_ppo_loss -> Engine.backward
      -> one rank:       sc_adam_flat
         data parallel:  sc_dp_barrier -> sc_dp_reduce_adam -> sc_dp_barrier   (peer.PeerReduceAdam: reduce-scatter +
                         sharded Adam + all-gather in one pass over NVLink peer memory; default when the policy's buckets
                         are in peer memory)
         or              NCCL all-reduce of the flat fp32 bucket -> sc_adam_flat   (fused_dp=False / SC_DP_FUSED=0;
                         SC_DP_OVERLAP=1 launches the all-reduce in three regions as backward finishes writing them)

Data parallelism: every rank holds whole trajectories; the only data-path
collectives are the flat-gradient all-reduce (one per optimizer step) and a 2-double all-reduce
of (sum A, sum A^2) per rollout for the global advantage normalisation.
"""
from __future__ import annotations

import os

import torch

from . import _lib, dp, ops
from .model import SemComPolicy
from .spec import SpecV0

F32 = torch.float32


class HostMinibatchFeeder:
    """Double-buffered pinned-host -> device staging of minibatches on a dedicated copy stream, so the
    H2D transfer of step i+1 overlaps the kernels of step i.  submit() enqueues the copies and returns
    (device tensors, ready event); the consumer stream must wait on the event before reading and call
    release() after its last use so the slot can be overwritten."""

    def __init__(self, device, slots: int = 2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.bufs = [dict() for _ in range(slots)]
        self.free_ev = [None] * slots
        self.i = 0

    def submit(self, host_mb: dict):
        k = self.i % len(self.bufs)
        self.i += 1
        buf = self.bufs[k]
        with torch.cuda.stream(self.stream):
            if self.free_ev[k] is not None:
                self.stream.wait_event(self.free_ev[k])          # consumer finished with this slot
            for name, src in host_mb.items():
                dst = buf.get(name)
                if dst is None or dst.shape != src.shape or dst.dtype != src.dtype:
                    dst = buf[name] = torch.empty(src.shape, dtype=src.dtype, device=self.device)
                dst.copy_(src, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        return buf, ready, k

    def release(self, k: int):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free_ev[k] = ev


class PPOAgent:
    def __init__(self, policy: SemComPolicy, spec: SpecV0 = SpecV0(), *, epochs: int = 1, minibatch: int = 4096,
                 snr_db: float = 10.0, channel: str = "awgn", seed: int = 1234, process_group=None,
                 fused_dp: bool | None = None):
        self.policy, self.spec = policy, spec
        self.epochs, self.minibatch = epochs, minibatch
        self.snr_db, self.channel, self.seed = snr_db, channel, seed
        self.step_count = 0
        self.gen = torch.Generator().manual_seed(seed)
        self.pg = process_group
        self.world, self.rank = dp.world_info(process_group)
        dev = policy.device
        # a new agent is a new optimizer (the oracle builds a fresh torch.optim.Adam): the flat Adam moments live next to
        # the parameters in the policy's FlatParams, so they are reset here; checkpoint.load() restores them afterwards
        policy.flat.m.zero_()
        policy.flat.v.zero_()
        self.stats = torch.zeros(8, device=dev, dtype=F32)      # pi, v, ent, kl, clipfrac, l_sem, -, -
        self.sums = torch.zeros(2, device=dev, dtype=torch.float64)
        # device-side mirror of step_count: the channel seed and Adam's bias correction read it inside the kernels,
        # so a captured CUDA graph of the step advances both on every replay.  sc_step_begin (the first kernel of a
        # step) increments it, so it equals step_count between steps and step_count + 1 inside one.
        self.dev_step = torch.zeros(1, device=dev, dtype=torch.int64)
        self._graphs = {}                                        # key (shapes + every baked-in scalar) -> (CUDAGraph, static inputs)
        self._overlap = None                                     # dp.OverlappedGradAllReduce, built on first use
        # opt-in: measured 2 % SLOWER at 2 GPUs (21.6 -> 22.0-22.3 ms): the NCCL kernels take SMs away from the
        # persistent one-CTA-per-SM GEMMs, whose displaced CTAs then run as a second wave (profiles/README.md)
        self.overlap_allreduce = os.environ.get("SC_DP_OVERLAP", "0") != "0"
        # data parallel: gradient exchange fused with the optimizer over peer memory (default whenever the policy's flat
        # buckets are in peer memory, i.e. it was built under an initialised NCCL process group); construction is a collective
        auto = fused_dp is None
        if auto:
            fused_dp = self.world > 1 and policy.flat.peer and not self.overlap_allreduce
        self._fused = None
        if fused_dp and self.world > 1:
            from .peer import PeerReduceAdam
            cache = policy.flat.__dict__.setdefault("_peer_reducers", {})    # one mapping per (buckets, group): a second
            if id(process_group) not in cache:                               # agent on the same policy reuses it
                try:
                    cache[id(process_group)] = PeerReduceAdam(policy.flat, process_group)
                except _lib.SemcomError as e:
                    # raised on every rank together (the ranks agree on the failure): GPUs without peer access / IPC.
                    # An explicit fused_dp=True propagates; the default takes the library's other schedule (NCCL all-reduce
                    # + sc_adam_flat, both on the GPU) and says so.
                    if not auto:
                        raise
                    import warnings
                    warnings.warn(f"peer-memory optimizer step unavailable, using the NCCL all-reduce schedule: {e}")
                    cache[id(process_group)] = None
            self._fused = cache[id(process_group)]
        self._lsem_weight = 1.0 / self.world                     # this rank's share of the global-minibatch mean L_sem
        self.last_perms = []                                     # permutations / per-minibatch statistics of the last
        self.last_step_stats = []                                #   update() (parity tests)
        self.allreduce_events = None                             # (start, end) CUDA events around the last gradient all-reduce
        self.replayed_launches = 0                               # library kernels executed through CUDA-graph replays (the
                                                                 # library's own counter only sees launches at capture time)

    # ---- one optimizer step on one (local) minibatch ---------------------------------------------
    def minibatch_step(self, cam, bev, actions, logp_old, adv, values_unused, ret, sample_offset: int = 0,
                       sync_stats: bool = True, graph: bool = False, *, row_idx=None, sample_idx=None,
                       global_batch: int | None = None, arena_cap: int = 0, obs_gathered: bool = False):
        """cam/bev fp32 or bf16, actions int64, logp_old/adv/ret fp32 — all on the policy's device.
        With world_size > 1 the arguments are this rank's shard of the global minibatch and
        `sample_offset` is its first global sample index (noise counters are global).
        graph=True: the step's ~130 launches are captured once per batch size in a CUDA graph and replayed (+2.5 % on one
        B200: no launch gaps); the arguments are copied into the graph's static input buffers.  With world_size > 1 the
        NCCL all-reduce of the flat gradient bucket is captured with the step (the single-all-reduce schedule only).  graph="inplace": one graph per SET of
        input buffers (keyed by their addresses), replayed without any copy — for callers that cycle through a few
        persistent device buffers (HostMinibatchFeeder slots); the caller keeps those buffers alive.
        row_idx (int64 [B], device): the minibatch is rows row_idx of the given tensors (whole-rollout storage) — the
        gather is fused into the first kernels (sc_patchify_bf16, sc_ppo_loss), no gathered copy is made.
        obs_gathered=True: cam / bev already hold exactly the B rows of this minibatch (update() streams them from pinned
        host memory with sc_gather_rows); row_idx then only indexes actions / logp_old / adv / ret.
        sample_idx (int64 [B], device) + global_batch: this rank's samples sit at positions sample_idx of a global
        minibatch of `global_batch` samples (ragged data-parallel slices of update())."""
        if graph and row_idx is None and sample_idx is None and (self.world == 1 or not self.overlap_allreduce):
            return self._graphed_step(cam, bev, actions, logp_old, adv, ret, sample_offset, sync_stats,
                                      inplace=(graph == "inplace"))
        self._step_body(cam, bev, actions, logp_old, adv, ret, sample_offset, device_counters=False, row_idx=row_idx,
                        sample_idx=sample_idx, global_batch=global_batch, arena_cap=arena_cap, obs_gathered=obs_gathered)
        if not sync_stats:
            return None
        return self.read_stats()

    def _step_body(self, cam, bev, actions, logp_old, adv, ret, sample_offset, device_counters: bool, row_idx=None,
                   sample_idx=None, global_batch=None, arena_cap: int = 0, obs_gathered: bool = False):
        """the kernel schedule of one optimizer step.  device_counters=True: seed and Adam step come from
        self.dev_step inside the kernels (graph capture / replay), otherwise from the host's step_count."""
        sp, pol = self.spec, self.policy
        eng, flat = pol.engine, pol.flat
        Bl = cam.shape[0] if row_idx is None else row_idx.numel()
        gb = Bl * self.world if global_batch is None else int(global_batch)
        # zero the gradient bucket and the statistics, dev_step += 1 (one launch)
        ops.step_begin(flat.g32, self.stats, self.dev_step)
        # dev_step already counts this step: seed + step_count = (seed - 1) + dev_step, Adam step = 0 + dev_step
        seed = self.seed - 1 if device_counters else self.seed + self.step_count
        sdev = self.dev_step if device_counters else None
        # per-sample loss terms are means over the GLOBAL minibatch; l_sem is a local mean weighted by this rank's share
        inv_m, sem_coef, self._lsem_weight = dp.loss_scales(Bl, self.world, sp.c_sem, global_batch=gb)
        if self.world > 1 and self._overlap is None:
            self._overlap = dp.OverlappedGradAllReduce(flat.g32, flat.offsets, self.pg)
        ov = self._overlap if (self.world > 1 and self.overlap_allreduce) else None
        if Bl > 0:
            arena = eng.arena(Bl, cap=arena_cap) if arena_cap else None
            if obs_gathered:
                assert cam.shape[0] == Bl and bev.shape[0] == Bl, (cam.shape, bev.shape, Bl)
            ho, A = eng.forward(cam, bev, self.snr_db, self.channel, seed, None, None, sample_offset, train=True,
                                seed_dev=sdev, row_idx=None if obs_gathered else row_idx, sample_idx=sample_idx, arena=arena)
            eng.l_sem(A, Bl, self.stats[5:6], grad_coef=sem_coef)
            dho = A.buf("ac.dho", tuple(ho.shape), F32)
            ops.ppo_loss(ho, actions, logp_old, adv, ret, sp.n_actions, sp.clip_eps, sp.c_v, sp.c_ent, inv_m,
                         dho=dho, stats=self.stats, idx=row_idx)
            # data parallel, overlap mode: the decoder's gradients are all-reduced while the encoders' backward still
            # runs, the camera encoder's while the BEV encoder's runs (three contiguous regions, dp.grad_buckets)
            eng.backward(A, Bl, dho, None, self.channel, self.snr_db, seed, None, sample_offset, seed_dev=sdev,
                         grad_ready=ov.ready if ov is not None else None, sample_idx=sample_idx)
        elif ov is not None:                                     # a rank without samples still joins every collective
            for name in ("dec", "cam_enc", "bev_enc"):
                ov.ready(name)
        timed = self.world > 1 and not torch.cuda.is_current_stream_capturing()    # timing events cannot be captured
        if timed:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        if self._fused is not None:
            # barrier -> reduce-scatter + sharded Adam + all-gather over peer memory -> barrier (three launches)
            if device_counters:
                self._fused.step(sp.lr, sp.beta1, sp.beta2, sp.adam_eps, 0, step_dev=self.dev_step)
            else:
                self._fused.step(sp.lr, sp.beta1, sp.beta2, sp.adam_eps, self.step_count + 1)
                self.step_count += 1
            if timed:
                ev1.record()
                self.allreduce_events = (ev0, ev1)
            return
        if ov is not None:
            ov.finish()
        else:
            dp.allreduce_flat_grads(flat.g32, self.pg)
        if timed:
            ev1.record()
            self.allreduce_events = (ev0, ev1)
        if device_counters:
            ops.adam_flat(flat.p32, flat.g32, flat.m, flat.v, flat.p16, sp.lr, sp.beta1, sp.beta2, sp.adam_eps,
                          0, grad_scale=1.0, step_dev=self.dev_step)
        else:
            ops.adam_flat(flat.p32, flat.g32, flat.m, flat.v, flat.p16, sp.lr, sp.beta1, sp.beta2, sp.adam_eps,
                          self.step_count + 1, grad_scale=1.0)
            self.step_count += 1

    def _graph_key(self, cam, bev, sample_offset):
        """everything a captured step bakes in: shapes / dtypes, the noise offset, and every scalar the kernels receive
        by value (channel, SNR, Rician K, the whole spec incl. lr and the loss coefficients, seed)"""
        snr = ("ptr", self.snr_db.data_ptr()) if torch.is_tensor(self.snr_db) else float(self.snr_db)
        return (tuple(cam.shape), cam.dtype, tuple(bev.shape), bev.dtype, int(sample_offset), snr, self.channel,
                float(self.policy.engine.rician_k), self.spec, int(self.seed))

    def _graphed_step(self, cam, bev, actions, logp_old, adv, ret, sample_offset, sync_stats, inplace=False):
        key = self._graph_key(cam, bev, sample_offset)
        if inplace:
            key = key + tuple(t.data_ptr() for t in (cam, bev, actions, logp_old, adv, ret))
        entry = self._graphs.get(key)
        if entry is None:
            # the first step of this shape runs eagerly (allocates the activation arena, sets kernel attributes);
            # then the same schedule is recorded, not executed, against static copies of the inputs
            self._step_body(cam, bev, actions, logp_old, adv, ret, sample_offset, device_counters=False)
            static = [t if inplace else t.clone() for t in (cam, bev, actions, logp_old, adv, ret)]
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count()
            with torch.cuda.graph(g):
                self._step_body(*static, sample_offset, device_counters=True)
            self._graphs[key] = (g, static, _lib.launch_count() - l0)
        else:
            g, static, n_launches = entry
            self.replayed_launches += n_launches
            for dst, src in zip(static, (cam, bev, actions, logp_old, adv, ret)):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
            g.replay()
            self.step_count += 1
        if not sync_stats:
            return None
        return self.read_stats()

    def _stats_snapshot(self) -> torch.Tensor:
        """device copy of this step's statistics, this rank's L_sem already weighted by its share of the global minibatch
        (so that a SUM over ranks gives the global-minibatch means); no host synchronisation"""
        st = self.stats.clone()
        if self.world > 1:
            st[5] *= self._lsem_weight
        return st

    def _stats_dict(self, s) -> dict:
        out = {"pi": s[0], "v": s[1], "ent": s[2], "kl": s[3], "clipfrac": s[4], "l_sem": s[5]}
        sp = self.spec
        out["total"] = out["pi"] + sp.c_v * out["v"] - sp.c_ent * out["ent"] + sp.c_sem * out["l_sem"]
        return out

    def read_stats(self) -> dict:
        st = self._stats_snapshot()
        if self.world > 1:
            torch.distributed.all_reduce(st, group=self.pg)
        return self._stats_dict(st.cpu().tolist())                # the step's device->host read

    def allreduce_ms(self):
        """device time of the last gradient all-reduce (None on one rank); synchronises"""
        if self.allreduce_events is None:
            return None
        ev0, ev1 = self.allreduce_events
        ev1.synchronize()
        return ev0.elapsed_time(ev1)

    # ---- advantages ----------------------------------------------------------------------------------
    def compute_advantages(self, rewards, values, dones, last_value):
        sp = self.spec
        self.sums.zero_()
        adv, ret = ops.gae(rewards, values, dones, last_value, sp.gamma, sp.lam, sums=self.sums)
        _, count = dp.global_adv_moments(self.sums, adv.numel(), self.pg)
        ops.adv_normalize(adv, self.sums, count, sp.adv_eps)
        return adv, ret

    def _update_streamed(self, steps, cam_host, bev_host, act, lp, adv, ret, mb, cap, snaps):
        """update()'s optimizer steps with the observations streamed from pinned host memory: the rows of step j + 1 are
        gathered over PCIe (sc_gather_rows, side stream, 16 CTAs) into the other staging slot while step j
        runs; events order slot reuse in both directions."""
        dev = self.policy.device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_obs_stream", None) is None:
            self._obs_stream, self._obs_stage = torch.cuda.Stream(dev), {}
        cs = self._obs_stream
        rows_max = max(nr for _, _, nr in steps)
        key = (rows_max, cam_host.dtype, tuple(cam_host.shape[1:]), bev_host.dtype, tuple(bev_host.shape[1:]))
        if self._obs_stage.get("key") != key:
            self._obs_stage = {"key": key, "slots": [
                (torch.empty((rows_max, *cam_host.shape[1:]), dtype=cam_host.dtype, device=dev),
                 torch.empty((rows_max, *bev_host.shape[1:]), dtype=bev_host.dtype, device=dev)) for _ in range(2)]}
        slots = self._obs_stage["slots"]
        idx_ready = torch.cuda.Event()
        idx_ready.record(main)                                    # the index uploads were enqueued on this stream
        cs.wait_event(idx_ready)
        ready, free = [None, None], [None, None]

        def prefetch(j):
            rows_d, _, nr = steps[j]
            k = j & 1
            with torch.cuda.stream(cs):
                if free[k] is not None:
                    cs.wait_event(free[k])                        # the step that read this slot has finished with it
                if nr > 0:
                    # 16 CTAs saturate PCIe; more would only displace more GEMM CTAs of the step running beside it
                    ops.gather_rows(cam_host, rows_d, slots[k][0][:nr], max_ctas=16)
                    ops.gather_rows(bev_host, rows_d, slots[k][1][:nr], max_ctas=16)
                ready[k] = torch.cuda.Event()
                ready[k].record(cs)

        prefetch(0)
        for j, (rows_d, pos_d, nr) in enumerate(steps):
            k = j & 1
            if j + 1 < len(steps):
                prefetch(j + 1)                                   # queued before this step's launches: runs beside them
            main.wait_event(ready[k])
            self.minibatch_step(slots[k][0][:nr], slots[k][1][:nr], act, lp, adv, None, ret, sample_offset=0,
                                row_idx=rows_d, sample_idx=pos_d, global_batch=mb, arena_cap=cap, sync_stats=False,
                                obs_gathered=True)
            free[k] = torch.cuda.Event()
            free[k].record(main)
            snaps.append(self._stats_snapshot())

    # ---- full update over a rollout ----------------------------------------------------------------
    def update(self, rollout: dict) -> dict:
        """rollout: this rank's trajectories (whole environments; dict of [T, N_local, ...] tensors, host or device).

        Matches `oracle/ppo.py: PPOAgent.update` step for step: the permutation of every epoch is the oracle's
        `torch.randperm(T * N_global, generator=cpu_gen(seed))` (bit-exact, host RNG).  With world_size > 1 every rank
        draws the SAME global permutation and processes, of each global minibatch, the samples whose environment it
        owns, at their global positions for the noise counters — so the N-rank update equals the 1-rank update on the
        concatenated rollout (tests/test_dp_gloo.py, scripts/gpu_dp_check.py).
        Device-resident or pageable observations go to the device once and minibatches are never materialised: the gather
        by `perm` is fused into sc_patchify_bf16 / sc_ppo_loss (row-index argument).  Observations in pinned host memory are
        streamed instead: each minibatch's rows cross PCIe (sc_gather_rows) while the previous step computes.
        Statistics are read back once, after the last step."""
        dev = self.policy.device
        to = lambda x, dt=F32: x.to(dev, dt, non_blocking=True).contiguous()
        rewards, values, dones = to(rollout["rewards"]), to(rollout["values"]), to(rollout["dones"])
        T, N = rewards.shape
        adv, ret = self.compute_advantages(rewards, values, dones, to(rollout["last_value"]))
        n = T * N
        obs_dt = lambda x: x.dtype if x.dtype == torch.bfloat16 else F32
        # Observations in PINNED host memory are not copied wholesale: every minibatch's rows are pulled over PCIe by
        # sc_gather_rows on a side stream while the previous optimizer step computes (double-buffered staging), so the
        # transfer of the rollout (872 MB for 8192 samples, ~16 ms of PCIe) leaves the critical path except for the
        # first minibatch.  SC_UPDATE_STREAM=0, pageable or device-resident rollouts: one copy up front, the gather fused
        # into sc_patchify_bf16.
        stream_obs = (os.environ.get("SC_UPDATE_STREAM", "1") != "0" and
                      all((not rollout[k].is_cuda) and rollout[k].is_pinned() and rollout[k].is_contiguous() and
                          rollout[k].dtype in (torch.bfloat16, F32) for k in ("cam", "bev")))
        if stream_obs:
            cam_h = rollout["cam"].view(n, *rollout["cam"].shape[2:])
            bev_h = rollout["bev"].view(n, *rollout["bev"].shape[2:])
        else:
            cam_h = to(rollout["cam"], obs_dt(rollout["cam"])).reshape(n, *rollout["cam"].shape[2:])
            bev_h = to(rollout["bev"], obs_dt(rollout["bev"])).reshape(n, *rollout["bev"].shape[2:])
        act = to(rollout["actions"], torch.int64).reshape(n)
        lp = to(rollout["logp"]).reshape(n)
        adv, ret = adv.reshape(n), ret.reshape(n)
        env_lo, n_envs_global = dp.env_shard(N, self.pg)
        n_global = T * n_envs_global
        mb = min(self.minibatch, n_global)                         # `minibatch` is the GLOBAL minibatch on any world size
        tot, cnt, snaps = {}, 0, []
        self.last_perms, self.last_step_stats = [], []
        steps, cap = [], 0
        for _ in range(self.epochs):
            perm = torch.randperm(n_global, generator=self.gen)    # host, seeded: bit-exact with the oracle
            self.last_perms.append(perm)
            slices = [dp.local_minibatch_indices(perm[i:i + mb], n_envs_global, env_lo, N)
                      for i in range(0, n_global - mb + 1, mb)]
            # one capacity arena for the ragged local slices of the update (exact-size arena on one rank)
            if self.world > 1:
                cap = max(cap, (max(int(r.numel()) for r, _ in slices) + 255) // 256 * 256)
            # the index vectors go to the device up front and the per-minibatch statistics stay there until the update is
            # over: no host round trip between optimizer steps (a per-step read leaves the GPU idle while the host
            # enqueues the next step's ~130 launches)
            steps += [(rows.to(dev, non_blocking=True), None if self.world == 1 else pos.to(dev, non_blocking=True),
                       int(rows.numel())) for rows, pos in slices]
        if stream_obs and steps:
            self._update_streamed(steps, cam_h, bev_h, act, lp, adv, ret, mb, cap, snaps)
        else:
            for rows_d, pos_d, _ in steps:
                self.minibatch_step(cam_h, bev_h, act, lp, adv, None, ret, sample_offset=0, row_idx=rows_d,
                                    sample_idx=pos_d, global_batch=mb, arena_cap=cap, sync_stats=False)
                snaps.append(self._stats_snapshot())
        if snaps:
            S = torch.stack(snaps)
            if self.world > 1:
                torch.distributed.all_reduce(S, group=self.pg)    # one small all-reduce per update, not one per step
            for row in S.cpu().tolist():                          # the update's only device->host read
                s = self._stats_dict(row)
                self.last_step_stats.append(s)
                for k, v in s.items():
                    tot[k] = tot.get(k, 0.0) + v
                cnt += 1
        return {k: v / max(cnt, 1) for k, v in tot.items()}
