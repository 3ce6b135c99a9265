"""Same-process A/B of sc_debug_set switches on the full PPO step (same box, same clocks, interleaved repeats).

    python scripts/gpu_ab_debug.py [--steps 10] [--detail] base: pf_off:gemm_dbg=256,mlp_dbg=1 ...

Every argument is `name:key=int,key=int,...`; the keys not named by a configuration are reset to 0 (pdl to its default
off).  Prints ms/step per repeat and a per-kernel table of the profile pass; writes gpurun_out/ab_debug.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_sc_b200 import ops, synthetic  # noqa: E402
from multimodal_sc_b200.model import SemComPolicy  # noqa: E402
from multimodal_sc_b200.ppo import PPOAgent  # noqa: E402
from multimodal_sc_b200.spec import SpecV0  # noqa: E402

args = sys.argv[1:]
steps, detail = 10, False
if "--steps" in args:
    i = args.index("--steps")
    steps = int(args[i + 1])
    del args[i:i + 2]
if "--detail" in args:
    detail = True
    args.remove("--detail")
configs = []
for a in args:
    name, _, kv = a.partition(":")
    configs.append((name, dict((k, int(v)) for k, v in (x.split("=") for x in kv.split(",") if x))))
all_keys = sorted({k for _, d in configs for k in d})

B = 4096
dev = torch.device("cuda", 0)
sp = SpecV0()
pol = SemComPolicy(sp, device=dev)
ag = PPOAgent(pol, sp, minibatch=B, snr_db=10.0, channel="awgn", seed=1234)
mbs = [{k: v.to(dev) for k, v in synthetic.minibatch(sp, B, seed=100 + i, obs_dtype=torch.bfloat16).items()} for i in range(2)]


def step(i):
    m = mbs[i & 1]
    return ag.minibatch_step(m["cam"], m["bev"], m["actions"], m["logp_old"], m["adv"], None, m["ret"], sync_stats=False)


def configure(d):
    for k in all_keys:
        ops.debug_set(k, d.get(k, 0))


def timed(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# numerics: one step from identical state under each configuration -> same statistics, same gradients up to the
# order of the fp32 atomics
state = {k: v.clone() for k, v in pol.state_dict().items()}
ref, checks = None, {}
for name, d in configs:
    configure(d)
    pol.load_state_dict(state)
    pol.flat.m.zero_()
    pol.flat.v.zero_()
    ag.step_count = 0
    ag.dev_step.zero_()
    step(0)
    torch.cuda.synchronize()
    g, st = pol.flat.g32.clone(), ag.stats.clone()
    if ref is None:
        ref = (g, st)
    checks[name] = {"grad_max_abs_diff": float((g - ref[0]).abs().max()), "grad_max_abs": float(ref[0].abs().max()),
                    "stats_max_abs_diff": float((st - ref[1]).abs().max())}
print("checks:", json.dumps(checks))

res, tables = {}, {}
for rep in range(3):
    for name, d in configs:
        configure(d)
        for i in range(3):
            step(i)
        res.setdefault(name, []).append(round(timed(steps), 3))
for name, d in configs:
    configure(d)
    step(0)
    prof = ops.profile()
    prof.detail = detail
    with prof:
        step(1)
    agg = {}
    for n, ms, fl, by in prof.records:
        e = agg.setdefault(n, [0, 0.0])
        e[0] += 1
        e[1] += ms
    tables[name] = {k: [v[0], round(v[1], 4)] for k, v in agg.items()}
names = [n for n, _ in configs]
print("ms/step:", json.dumps(res))
keys = sorted(tables[names[0]], key=lambda k: -tables[names[0]][k][1])
print(f"{'kernel':36s} n  " + "  ".join(f"{n:>10s}" for n in names))
for k in keys:
    print(f"{k:36s} {tables[names[0]][k][0]:3d} " + "  ".join(f"{tables[n].get(k, [0, 0.0])[1]:10.3f}" for n in names))
print(f"{'total':36s}     " + "  ".join(f"{sum(v[1] for v in tables[n].values()):10.3f}" for n in names))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"steps": steps, "ms_per_step": res, "per_kernel_ms": tables}, open(os.path.join(ROOT, "gpurun_out", "ab_debug.json"), "w"), indent=1)
