"""torchrun --nproc-per-node 2 scripts/gpu_dp_check.py : the overlapped (bucketed, async) gradient all-reduce against
the single all-reduce after backward, same shards, same initial weights.  Prints the difference of the reduced
gradient buckets (should be summation-order noise of the local fp32 atomics) and the step times of both modes."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_sc_b200 import ops, synthetic  # noqa: E402
from multimodal_sc_b200.model import SemComPolicy  # noqa: E402
from multimodal_sc_b200.ppo import PPOAgent  # noqa: E402
from multimodal_sc_b200.spec import SpecV0  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sp = SpecV0()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mb = {k: v.to(dev) for k, v in synthetic.minibatch(sp, B, seed=100 + rank, obs_dtype=torch.bfloat16).items()}
grads = {}
times = {}
for mode in (False, True, False, True):
    pol = SemComPolicy(sp, device=dev, seed=0)
    ag = PPOAgent(pol, sp, minibatch=B, seed=5)
    ag.overlap_allreduce = mode
    # capture the reduced gradient bucket: Adam reads it last, so clone it through a wrapper around adam_flat
    orig = ops.adam_flat
    def spy(p32, g32, *a, **k):
        grads[mode] = g32.clone()
        return orig(p32, g32, *a, **k)
    ops.adam_flat = spy
    ag.minibatch_step(mb["cam"], mb["bev"], mb["actions"], mb["logp_old"], mb["adv"], None, mb["ret"], sample_offset=rank * B)
    ops.adam_flat = orig
    for _ in range(3):
        ag.minibatch_step(mb["cam"], mb["bev"], mb["actions"], mb["logp_old"], mb["adv"], None, mb["ret"], sample_offset=rank * B, sync_stats=False)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ag.minibatch_step(mb["cam"], mb["bev"], mb["actions"], mb["logp_old"], mb["adv"], None, mb["ret"], sample_offset=rank * B, sync_stats=False)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.setdefault("overlap" if mode else "single", []).append(round(float(t), 3))
d = (grads[True] - grads[False]).abs()
out = {"world": world, "grad_max_abs": float(grads[False].abs().max()), "grad_max_abs_diff": float(d.max()),
       "grad_mean_abs_diff": float(d.mean()), "ms_per_step": times}
if rank == 0:
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"dp_check_n{world}.json"), "w"), indent=1)
dist.destroy_process_group()
