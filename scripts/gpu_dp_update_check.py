"""torchrun --nproc-per-node 2 scripts/gpu_dp_update_check.py

SURVEY.md §8e at update() level on real GPUs: a 2-rank PPOAgent.update() — every rank owns an UNEQUAL run of whole
environments (3 and 5 of 8), draws the same global permutation and processes the samples it owns of every global minibatch
(ragged local slices, capacity arena, global noise positions, one NCCL all-reduce of the flat gradient bucket per step) —
against the 1-rank update() on the concatenated rollout (run on every rank in a 1-rank process group, same seeds).

Checks: identical permutations; per-minibatch statistics; final parameters.  Both runs are the native bf16 path, per-sample
arithmetic does not depend on the batch composition, so the only differences are fp32 summation order (split-K red.add,
NCCL ring order) — far below the bf16 tolerance.  Writes gpurun_out/dp_update_check_n<world>.json."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_sc_b200 import synthetic  # noqa: E402
from multimodal_sc_b200.model import SemComPolicy  # noqa: E402
from multimodal_sc_b200.ppo import PPOAgent  # noqa: E402
from multimodal_sc_b200.spec import SpecV0  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
solo = [dist.new_group([r]) for r in range(world)][rank]           # 1-rank group for the reference run
sp = SpecV0()
T, N, MB, EPOCHS = 16, 8, 32, 2
ro = synthetic.rollout(sp, T, N, seed=21)                           # identical on every rank
# unequal contiguous environment shards
cuts = [0, 3, 8] if world == 2 else [round(i * N / world) for i in range(world + 1)]
lo, hi = cuts[rank], cuts[rank + 1]
local_ro = {k: (v[:, lo:hi] if v.dim() >= 2 else v[lo:hi]).contiguous() for k, v in ro.items()}

out = {"world": world, "T": T, "N": N, "global_minibatch": MB, "epochs": EPOCHS, "env_shards": cuts}
for channel in ("awgn", "rayleigh"):
    pol_dp = SemComPolicy(sp, device=dev, seed=0)
    ag_dp = PPOAgent(pol_dp, sp, minibatch=MB, epochs=EPOCHS, seed=7, channel=channel, snr_db=5.0)
    # the Rayleigh pass feeds the data-parallel update from PINNED host memory: observation rows are then streamed per minibatch
    # (sc_gather_rows, one step ahead) instead of copied up front — same result expected
    pinned = channel == "rayleigh"
    st_dp = ag_dp.update({k: v.pin_memory() for k, v in local_ro.items()} if pinned else local_ro)
    assert (getattr(ag_dp, "_obs_stream", None) is not None) == pinned
    pol_1 = SemComPolicy(sp, device=dev, seed=0)
    ag_1 = PPOAgent(pol_1, sp, minibatch=MB, epochs=EPOCHS, seed=7, channel=channel, snr_db=5.0, process_group=solo)
    assert ag_1.world == 1 and ag_dp.world == world
    st_1 = ag_1.update(ro)
    perms_equal = all(torch.equal(a, b) for a, b in zip(ag_dp.last_perms, ag_1.last_perms))
    local_counts = None
    worst, first = {}, {}
    for i, (a, b) in enumerate(zip(ag_1.last_step_stats, ag_dp.last_step_stats)):
        for k in ("pi", "v", "ent", "kl", "l_sem", "total"):
            d = abs(a[k] - b[k]) / max(abs(a[k]), 0.05)
            worst[k] = max(worst.get(k, 0.0), d)
            if i == 0:
                first[k] = d
    dpar = (pol_dp.flat.p32 - pol_1.flat.p32).abs()
    # every rank must hold the same parameters after the update
    chk = pol_dp.flat.p32.double().sum().reshape(1)
    both = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(both, chk)
    res = {"perms_equal": bool(perms_equal), "steps": len(ag_dp.last_step_stats), "dp_rollout_in_pinned_host_memory": pinned,
           "dp_optimizer_step": "peer memory" if ag_dp._fused is not None else "nccl",
           "first_step_rel_stat_diff": first, "worst_rel_stat_diff": worst, "update_mean_stats_dp": st_dp, "update_mean_stats_1rank": st_1,
           "param_max_abs_diff": float(dpar.max()), "param_frac_diff_gt_0.02lr": float((dpar > 0.02 * sp.lr).float().mean()),
           "lr": sp.lr, "ranks_hold_identical_params": bool(all(float(b) == float(both[0]) for b in both))}
    # step 0 runs on identical weights: only fp32 summation order differs.  Later steps: an element whose gradient is
    # within summation-order noise of zero moves by +-lr either way (Adam's first steps are sign-like), which perturbs
    # the following steps' statistics at the 1e-3 level
    res["ok"] = bool(perms_equal and max(first.values()) < 1e-4 and max(worst.values()) < 2e-2
                     and res["ranks_hold_identical_params"])
    out[channel] = res
if rank == 0:
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"dp_update_check_n{world}.json"), "w"), indent=1)
ok = all(out[c]["ok"] for c in ("awgn", "rayleigh"))
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
sys.stderr.flush()
os._exit(0 if ok else 1)        # no collective destroy: peer-memory mappings and communicators go with the process
