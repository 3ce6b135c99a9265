#!/usr/bin/env python
"""bench.py — headline benchmark of the SemCom-PPO hot path (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W            # native sm_100a arm
    python bench.py --impl reference --steps K --warmup W    # the oracle on the host CPU

Workload (config.workload): cfg3/cfg5 of BASELINE.json — one PPO optimizer step
(enc -> channel -> dec -> policy forward, fused PPO loss, full backward, Adam) on a
minibatch of 4096 samples PER GPU, bf16 tensor-core GEMMs, AWGN 10 dB, synthetic data.
At N GPUs the global batch is 4096*N (N=8 -> 32768 = cfg5), data-parallel with one NCCL
all-reduce of the flat fp32 gradient bucket per step: weak scaling.

Prints ONE JSON line (rank 0).  `value` = samples/s with inputs resident in HBM, `e2e` = the same
step driven through the public API (PPOAgent.minibatch_step) from pinned HOST buffers with the
H2D copy of every step's inputs and the D2H read of its statistics inside the timed region.

Beside the contract keys the line carries: `roofline` (dominant kernel family, both views, per-kernel
table with algorithmic GB/s for every kernel of the step), `cpu_baseline`, `update_e2e` (the WHOLE
PPOAgent.update() from a pinned host rollout: H2D of the rollout, fused GAE scan + advantage
normalisation, host permutation, minibatch gather fused into the first kernels, every minibatch step,
D2H of the statistics), `allreduce_ms` (N > 1: device time of the gradient all-reduce, the scaling
limiter) and `secondary` (BASELINE.json's other configs measured in the same run: cfg2 forward eager +
CUDA graph, cfg4 Rayleigh sweep per SNR, cfg1 the oracle's own 1-env CPU update).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "samples/sec enc->channel->dec->policy fwd+bwd"
UNIT = "samples/s"
MB_PER_GPU = 4096
SNR_DB = 10.0


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i] == "Active" for r in self.rows)]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_oracle_cfg1():
    """cfg1 (BASELINE.json configs[0]): the oracle's CPU entry point at its smallest config — 1 environment, T = 128,
    one update() (GAE + 1 epoch over one minibatch of 128), fp32.  -> dict"""
    import torch
    from multimodal_sc_b200.spec import SpecV0
    from multimodal_sc_b200 import synthetic
    from oracle.model import SemComPolicy
    from oracle.ppo import PPOAgent
    torch.set_num_threads(os.cpu_count() or 1)
    sp = SpecV0()
    ag = PPOAgent(SemComPolicy(sp), sp, minibatch=128, epochs=1, snr_db=SNR_DB, channel="awgn", seed=1234)
    ro = synthetic.rollout(sp, sp.rollout_t, 1, seed=0)
    t0 = time.perf_counter()
    st = ag.update(ro)
    dt = time.perf_counter() - t0
    return {"workload": "cfg1: oracle update(), 1 env x T=128, minibatch 128, fp32, host CPU", "ms": dt * 1e3,
            "samples_per_s": sp.rollout_t / dt, "cores": torch.get_num_threads(), "total_loss": st["total"]}


def cpu_oracle_rate(steps: int, warmup: int, batch: int):
    """oracle PPO minibatch steps (fwd + loss + bwd + Adam), fp32, all host threads. -> (samples/s, ms/step, cores)"""
    import torch
    from multimodal_sc_b200.spec import SpecV0
    from multimodal_sc_b200 import synthetic
    from oracle.model import SemComPolicy
    from oracle.ppo import PPOAgent
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sp = SpecV0()
    pol = SemComPolicy(sp)
    ag = PPOAgent(pol, sp, minibatch=batch, snr_db=SNR_DB, channel="awgn", seed=1234)
    mb = synthetic.minibatch(sp, batch, seed=0)
    args = (mb["cam"], mb["bev"], mb["actions"], mb["logp_old"], mb["adv"], None, mb["ret"])
    for _ in range(warmup):
        ag.minibatch_step(*args)
    t0 = time.perf_counter()
    for _ in range(steps):
        ag.minibatch_step(*args)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 256
    steps = max(1, min(a.steps, 20))
    val, ms, cores = cpu_oracle_rate(steps, max(1, min(a.warmup, 2)), batch)
    sample = f"{steps} oracle PPO minibatch steps of {batch} samples (same shapes/seeds as the workload, fp32 eager PyTorch)"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "cfg3 PPO minibatch step on the host CPU (oracle port; upstream has no code)",
                       "minibatch": batch, "snr_db": SNR_DB, "channel": "awgn"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# secondary configs (BASELINE.json configs[0], [1], [3]) — measured in the same run, for the record
# --------------------------------------------------------------------------------------------------
def secondary_configs(pol, sp, torch, synthetic, _lib, ops, cpu_leg=True):
    eng = pol.engine
    out = {}

    def timeit(fn, n=30, warm=5):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    fwd_flops = None
    from multimodal_sc_b200.spec import fwd_flops_per_sample
    fwd_flops = fwd_flops_per_sample(sp)
    peaks = _peaks()
    # cfg2: forward only, B = 256, AWGN 10 dB — eager and as one CUDA-graph replay (device-side seed counter)
    B = 256
    mb = synthetic.minibatch(sp, B, seed=5, device=pol.device, obs_dtype=torch.bfloat16)
    l0 = _lib.launch_count()
    ms = timeit(lambda: eng.forward(mb["cam"], mb["bev"], SNR_DB, "awgn", 77, None, None, 0, train=False))
    c2 = {"workload": "cfg2: camera + LiDAR-BEV encoders -> AWGN 10 dB -> fusion decoder + heads forward, batch 256, bf16",
          "eager_ms": ms, "eager_samples_per_s": B / ms * 1e3, "launches_per_forward": (_lib.launch_count() - l0) // 35}
    ctr = torch.zeros(1, device=pol.device, dtype=torch.int64)

    def body():
        eng.forward(mb["cam"], mb["bev"], SNR_DB, "awgn", 77, None, None, 0, train=False, seed_dev=ctr)
        ctr.add_(1)

    body()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    ms = timeit(g.replay)
    c2.update({"graph_ms": ms, "graph_samples_per_s": B / ms * 1e3,
               "graph_frac_of_fwd_roofline": (B / ms * 1e3) * fwd_flops / (peaks["tf_sustained"] * 1e12)})
    out["cfg2"] = c2
    # cfg4: Rayleigh block fading, B = 1024, SNR sweep, forward + L_sem + backward of the semantic loss
    B = 1024
    mb = synthetic.minibatch(sp, B, seed=6, device=pol.device, obs_dtype=torch.bfloat16)
    lsem = torch.zeros(1, device=pol.device)
    dho = torch.zeros(B, sp.n_actions + 1, device=pol.device)
    per = {}
    for snr in (-5.0, 0.0, 5.0, 10.0, 15.0, 20.0):
        def step():
            ops.step_begin(pol.flat.g32, lsem, None)
            ho, A = eng.forward(mb["cam"], mb["bev"], snr, "rayleigh", 1234, None, None, 0, train=True)
            eng.l_sem(A, B, lsem, grad_coef=1.0)
            eng.backward(A, B, dho, None, "rayleigh", snr, 1234, None, 0)
        ms = timeit(step, n=10, warm=3)
        per[str(snr)] = {"ms": ms, "samples_per_s": B / ms * 1e3, "l_sem": float(lsem)}
    out["cfg4"] = {"workload": "cfg4: Rayleigh fading, SNR sweep -5..20 dB, semantic-distortion loss forward + backward, "
                               "batch 1024, bf16 (random-init weights: l_sem is not a trained value)", "per_snr_db": per}
    if cpu_leg:
        out["cfg1"] = cpu_oracle_cfg1()
    return out


# --------------------------------------------------------------------------------------------------
# native arm
# --------------------------------------------------------------------------------------------------
def run_native(a):
    import torch
    import torch.distributed as dist
    from multimodal_sc_b200 import _lib, ops, synthetic
    from multimodal_sc_b200.model import SemComPolicy
    from multimodal_sc_b200.ppo import HostMinibatchFeeder, PPOAgent
    from multimodal_sc_b200.spec import SpecV0, fwd_flops_per_sample

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    t_start = time.perf_counter()

    def phase(msg):                                   # progress marks on stderr (stdout carries the JSON line only)
        if rank == 0:
            print(f"[bench +{time.perf_counter() - t_start:6.1f}s] {msg}", file=sys.stderr, flush=True)

    def finish():
        """tear-down: captured graphs hold the NCCL communicator, so they go first; with N > 1 the process then leaves
        without the collective destroy (a rank that is already gone cannot answer it)"""
        ag._graphs.clear()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            os._exit(0)

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # torchrun pins OMP_NUM_THREADS=1: the synthetic host inputs (2 x 436 MB + a 872 MB rollout per rank) are generated
    # outside every timed region, with this rank's share of the host cores
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(world, 1)))
    # stdout carries exactly ONE JSON line: libraries that chat on fd 1 (NCCL prints its version there when
    # NCCL_DEBUG is set) are pointed at stderr until the line is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.require_device()
    for kv in a.debug_set:
        k, v = kv.split("=")
        ops.debug_set(k, int(v))
    sp = SpecV0()
    B = a.minibatch
    pol = SemComPolicy(sp, device=dev)
    ag = PPOAgent(pol, sp, minibatch=B, snr_db=SNR_DB, channel="awgn", seed=1234)
    # two distinct minibatches per rank (2 x 436 MB of bf16 observations >> 126 MB of L2)
    # observations: U[0,1) fp32 cast to bf16 once on the host, outside any timed region
    host = [synthetic.minibatch(sp, B, seed=100 + 2 * rank + i, pin=True, obs_dtype=torch.bfloat16) for i in range(2)]
    devb = [{k: v.to(dev) for k, v in h.items()} for h in host]

    def step_dev(i, sync_stats=False):
        m = devb[i & 1]
        return ag.minibatch_step(m["cam"], m["bev"], m["actions"], m["logp_old"], m["adv"], None, m["ret"],
                                 sample_offset=rank * B, sync_stats=sync_stats, graph=GRAPH)

    # the step is replayed as a CUDA graph (one graph per persistent input buffer set, no copies; with N > 1 the NCCL
    # all-reduce is captured with it); --eager launches kernel by kernel
    GRAPH = False if a.eager else "inplace"
    feeder = HostMinibatchFeeder(dev)
    pending = {}

    def step_host(i, last):
        """public-API step from pinned HOST buffers: this step's H2D was enqueued on the copy stream one
        step earlier (step 0: just now), so it overlaps the previous step's kernels; the stats read is a
        blocking D2H.  Exactly `steps` H2D transfers and `steps` D2H reads happen inside the timed region."""
        if i == 0:
            pending[0] = feeder.submit(host[0])
        m, ready, slot = pending.pop(i)
        if not last:
            pending[i + 1] = feeder.submit(host[(i + 1) & 1])
        torch.cuda.current_stream().wait_event(ready)
        st = ag.minibatch_step(m["cam"], m["bev"], m["actions"], m["logp_old"], m["adv"], None, m["ret"],
                               sample_offset=rank * B, sync_stats=True, graph=GRAPH)   # D2H read of the step's stats
        feeder.release(slot)
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    per_step = {}

    def timed(fn, steps, with_last=False, tag=None):
        """EXACTLY `steps` steps between barrier + synchronize on both sides, CUDA events, max over ranks; an event after
        every step additionally gives the per-step median."""
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        barrier()
        evs[0].record()
        for i in range(steps):
            fn(i, i == steps - 1) if with_last else fn(i)
            evs[i + 1].record()
        barrier()
        ms = torch.tensor([evs[0].elapsed_time(evs[-1])], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if tag:
            d = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(steps))
            per_step[tag] = {"median": d[len(d) // 2], "min": d[0], "max": d[-1]}
        return float(ms)

    phase("inputs ready, warm-up")
    for i in range(max(a.warmup, 3)):
        step_dev(i)
    phase("timed region (device-resident)")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count() + ag.replayed_launches
    ms_total = timed(step_dev, a.steps, tag="value")
    launches = _lib.launch_count() + ag.replayed_launches - l0   # eager launches + kernels inside the replayed graphs
    ar_ms = None
    if world > 1:                                             # device time of the gradient all-reduce: one eager step (the
        saved, GRAPH = GRAPH, False                           # graph replays carry no timing events), after the timed region
        step_dev(0)                                           # (the first eager step after replays re-primes NCCL's eager path)
        step_dev(1)
        ar_ms = ag.allreduce_ms()
        GRAPH = saved
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * a.steps / (ms_total * 1e-3)

    # end-to-end: host buffers, H2D + D2H inside the timed region
    phase("timed region (end to end from pinned host buffers)")
    for i in range(2):
        step_host(i, i == 1)
    ms_e2e = timed(step_host, a.steps, with_last=True, tag="e2e")
    e2e_val = world * B * a.steps / (ms_e2e * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    # roofline pass: one extra step with every launch bracketed by CUDA events on the launching stream
    phase("roofline pass")
    barrier()
    GRAPH = False                                   # per-launch events need the eager schedule
    with ops.profile() as prof:
        step_dev(0)
    summ = prof.summary()

    # whole update() from a pinned host rollout (GAE, permutation, fused gather, every minibatch step); at N > 1 each rank
    # owns N_ENVS whole environments and update() slices the global permutation by rank (ragged local minibatches)
    upd = None
    if not a.no_update:
        phase("update_e2e")
        # a rollout of `--update-steps` (default 4) global minibatches per GPU, built in chunks of <= 64 environments so that
        # the fp32 staging of the synthetic observations stays below 1.7 GB per rank
        T_ro, n_env = sp.rollout_t, max(a.update_steps * B // sp.rollout_t, 1)
        parts = []
        for c0 in range(0, n_env, 64):
            part = synthetic.rollout(sp, T_ro, min(64, n_env - c0), seed=300 + 16 * rank + c0 // 64)
            part["cam"], part["bev"] = part["cam"].bfloat16(), part["bev"].bfloat16()
            parts.append(part)
        ro = {k: torch.cat([p_[k] for p_ in parts], dim=0 if k == "last_value" else 1).contiguous().pin_memory()
              for k in parts[0]}
        del parts
        agu = PPOAgent(pol, sp, minibatch=B * world, epochs=1, snr_db=SNR_DB, channel="awgn", seed=4321,
                       process_group=None)
        agu.update(ro)                                          # warm-up (allocates the capacity arena on N > 1)
        reps = 3
        ms_u = timed(lambda i: agu.update(ro), reps)
        n_samples = T_ro * n_env * world
        upd = {"value": n_samples * reps / (ms_u * 1e-3), "unit": UNIT, "ms_per_update": ms_u / reps,
               "rollout": f"T={T_ro} x N={n_env} envs per GPU (bf16 observations, pinned host)",
               "minibatch_steps_per_update": (n_samples // (B * world)),
               "h2d_bytes_per_update": sum(v.numel() * v.element_size() for v in ro.values()),
               "includes": "H2D of the rollout (observation rows streamed per minibatch by sc_gather_rows, one step ahead), sc_gae + "
                           "sc_adv_normalize, host randperm, gather of the per-sample vectors fused into sc_ppo_loss, all minibatch "
                           "steps, one D2H of the statistics"}
        del ro
    if world > 1:
        dist.barrier()
    if rank != 0:
        finish()
        return
    phase("cpu baseline / secondary configs")
    peaks = _peaks()
    gemm = {k: v for k, v in summ.items() if k.startswith("gemm_")}
    g_ms = sum(v["ms"] for v in gemm.values())
    g_fl = sum(v["flops"] for v in gemm.values())
    g_n = sum(v["launches"] for v in gemm.values())
    step_ms_prof = sum(v["ms"] for v in summ.values())
    achieved = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    g_by = sum(v["bytes"] for v in gemm.values())
    # DRAM bytes per launch: NOT measured by this run (it needs ncu replay); taken from the committed `ncu --set full`
    # capture of this same command, whose file records the commit it was taken at (`traffic_source`)
    traffic, traffic_src = None, None
    for name in ("r2_gemm_dram_traffic.json", "r1_gemm_dram_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath) and B == MB_PER_GPU:
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_launch_avg")
            traffic_src = f"profiles/{name} (ncu capture at commit {tj.get('commit', 'unrecorded')})"
            break
    # which roof binds: arithmetic intensity of the launches (algorithmic FLOP / algorithmic bytes, each operand read
    # once, each result written once) against the ridge point of the measured peaks.  The D = 256 GEMMs of this model
    # sit below the ridge (AI ~150 < ~214 FLOP/B), so the attainable rate is AI x HBM bandwidth, not the tensor peak;
    # both views are reported and `bound` names the binding one.
    ai = g_fl / max(g_by, 1.0)
    ridge = peaks["tf_sustained"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    gbs = g_by / (g_ms * 1e-3) / 1e9 if g_ms > 0 else 0.0
    tensor_view = {"achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                   "frac": achieved / peaks["tf_sustained"]}
    hbm_view = {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"]}
    bound = "hbm" if ai < ridge else "tensor"
    main = hbm_view if bound == "hbm" else tensor_view
    roof = {"bound": bound, "kernel": "gemm_tcgen05_kernel (all operand-major variants)",
            "achieved": main["achieved"], "peak": main["peak"], "unit": main["unit"], "frac": main["frac"],
            "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": g_by / max(g_n, 1),
            "arithmetic_intensity_flop_per_byte": ai, "ridge_flop_per_byte": ridge,
            "tensor_view": tensor_view, "hbm_view": hbm_view,
            "peak_source": peaks["source"] + "; tensor view uses the sustained bf16 figure (kernels timed inside a long step)",
            "launches_per_step": g_n, "avg_launch_ms": g_ms / max(g_n, 1), "share_of_step": g_ms / max(step_ms_prof, 1e-9),
            "algorithmic_flops_per_step": g_fl,
            "by_kernel": {k: {"launches": v["launches"], "ms": round(v["ms"], 4),
                              "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1),
                              "gbs": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1)} for k, v in sorted(summ.items())}}
    # attainable roof per LAUNCH: max(algorithmic bytes / HBM copy bandwidth, FLOPs / sustained bf16 peak) — the family mixes
    # HBM-bound forward shapes with tensor-bound weight gradients, so each launch is compared with ITS binding roof
    def t_roof_ms(fl, by):
        return max(by / (peaks["hbm_gbs"] * 1e9), fl / (peaks["tf_sustained"] * 1e12)) * 1e3
    att = {}
    for n_, ms_, fl_, by_ in prof.records:
        d_ = att.setdefault(n_, [0.0, 0.0])
        d_[0] += ms_
        d_[1] += t_roof_ms(fl_, by_)
    g_att = [sum(v[i] for k, v in att.items() if k.startswith("gemm_")) for i in (0, 1)]
    s_att = [sum(v[i] for v in att.values()) for i in (0, 1)]
    roof["attainable"] = {"what": "sum over launches of max(bytes / HBM copy bandwidth, FLOPs / sustained bf16 peak) divided by the "
                                  "measured launch times: every launch against its own binding roof",
                          "gemm_family_frac": g_att[1] / max(g_att[0], 1e-9), "step_frac": s_att[1] / max(s_att[0], 1e-9),
                          "step_roof_ms": s_att[1], "step_measured_ms_launch_by_launch": s_att[0],
                          "by_kernel_frac": {k: round(v[1] / max(v[0], 1e-9), 3) for k, v in sorted(att.items())}}
    flops_step = 3.0 * fwd_flops_per_sample(sp) * B
    cpu = None
    secondary = None
    if world == 1 and not a.no_cpu:
        v, ms, cores = cpu_oracle_rate(8, 2, 256)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_step": ms,
               "sample": "8 timed (+2 warm-up) oracle PPO minibatch steps of 256 samples (fp32 eager PyTorch, all host threads)"}
    if world == 1 and not a.no_secondary:
        secondary = secondary_configs(pol, sp, torch, synthetic, _lib, ops, cpu_leg=not a.no_cpu)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "ms_per_step_median": per_step["value"]["median"], "ms_per_step_minmax": [per_step["value"]["min"], per_step["value"]["max"]],
            "config": {"workload": f"cfg3/cfg5: PPO minibatch step (fwd + fused loss + bwd + Adam), {B} samples per GPU, "
                                   f"AWGN 10 dB; global batch = {B} x {world} GPUs (weak scaling; 8 GPUs x 4096 = cfg5's 32768), "
                                   "DP all-reduce of the flat fp32 grads",
                       "minibatch_per_gpu": B, "global_batch": B * world, "snr_db": SNR_DB, "channel": "awgn",
                       "params": pol.flat.n_params(), "parallelism": f"dp{world}",
                       "schedule": "eager launches" if a.eager else "CUDA-graph replay of the whole step (PPOAgent.minibatch_step(graph='inplace'))",
                       "observations": "bf16 (fp32 U[0,1) cast once on the host before timing)",
                       "gradient_exchange": (None if world == 1 else
                                             ("sc_dp_barrier + sc_dp_reduce_adam + sc_dp_barrier over NVLink peer memory (reduce-scatter, "
                                              "sharded Adam, parameter all-gather)" if ag._fused is not None else
                                              "NCCL all-reduce of the flat fp32 bucket + sc_adam_flat")),
                       "l2_policy": "inputs larger than L2 (2 x 436 MB bf16 observation sets alternate; ~16 GB of "
                                    "activations stream per step)"},
            "model_tflops": flops_step / (ms_total / a.steps * 1e-3) / 1e12,
            "model_tflops_frac_of_sustained_peak": flops_step / (ms_total / a.steps * 1e-3) / 1e12 / peaks["tf_sustained"],
            "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": ms_e2e / a.steps, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 32},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof}
    line["e2e"]["ms_per_step_median"] = per_step["e2e"]["median"]
    if ar_ms is not None:
        line["allreduce_ms"] = ar_ms
        line["allreduce"] = {"ms": ar_ms, "bytes": pol.flat.numel * 4, "share_of_step": ar_ms / (ms_total / a.steps),
                             "impl": "peer" if ag._fused is not None else "nccl",
                             "what": ("CUDA events around sc_dp_barrier + sc_dp_reduce_adam + sc_dp_barrier (gradient reduce-scatter, "
                                      "sharded Adam and parameter all-gather over NVLink peer memory: the optimizer is INSIDE this time)"
                                      if ag._fused is not None else
                                      "CUDA events around the NCCL all-reduce of the flat fp32 gradient bucket") +
                                     " in one eager step right after the timed region, rank 0 (includes waiting for the slowest rank "
                                     "to finish its backward)"}
    if upd:
        line["update_e2e"] = upd
    if secondary:
        line["secondary"] = secondary
    if cpu:
        line["cpu_baseline"] = cpu
    if a.debug_set:
        line["config"]["debug"] = list(a.debug_set)      # an A/B run, not a headline number
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    phase("done")
    finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)      # >= 20 warm-up, >= 100 timed iterations
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--minibatch", type=int, default=MB_PER_GPU)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the cfg1 / cfg2 / cfg4 measurements")
    ap.add_argument("--no-update", action="store_true", help="skip the update_e2e measurement")
    ap.add_argument("--update-steps", type=int, default=4,
                    help="optimizer steps (global minibatches) per update() in the update_e2e measurement")
    ap.add_argument("--eager", action="store_true",
                    help="A/B: launch the step's kernels one by one instead of replaying the captured CUDA graph "
                         "(PPOAgent.minibatch_step(graph=\"inplace\"), the default: +2.5 %% at 1 GPU, +0.7 %% at 2)")
    ap.add_argument("--debug-set", action="append", default=[], metavar="KEY=INT",
                    help="A/B switch forwarded to sc_debug_set (e.g. nwg=1 slots=2); recorded in config.debug")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)


if __name__ == "__main__":
    main()
