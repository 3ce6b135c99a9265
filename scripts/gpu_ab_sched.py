"""Same-process A/B of the step schedule: programmatic dependent launch (sc_debug_set("pdl")) and the side stream
for the block wgrads (Engine.wgrad_stream).  usage: python scripts/gpu_ab_sched.py [steps] -> gpurun_out/ab_sched.json"""
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

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = 4096
dev = torch.device("cuda", 0)
sp = SpecV0()
pol = SemComPolicy(sp, device=dev)
ag = PPOAgent(pol, sp, minibatch=B, snr_db=10.0, channel="awgn", seed=1234)
mbs = [{k: v.to(dev) for k, v in synthetic.minibatch(sp, B, seed=100 + i, obs_dtype=torch.bfloat16).items()} for i in range(2)]
side_stream = torch.cuda.Stream(dev)


def step(i, graph=False):
    m = mbs[i & 1]
    return ag.minibatch_step(m["cam"], m["bev"], m["actions"], m["logp_old"], m["adv"], None, m["ret"], sync_stats=False,
                             graph=graph)


def configure(pdl, side):
    ops.debug_set("pdl", int(pdl))
    pol.engine.wgrad_stream = side_stream if side else None


def timed(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# numerics: one step from identical state under each schedule -> identical statistics and (up to the order of the
# fp32 atomics) identical gradients
state = {k: v.clone() for k, v in pol.state_dict().items()}
ref = None
checks = {}
for name, (pdl, side) in {"base": (0, 0), "pdl": (1, 0), "side": (0, 1), "pdl+side": (1, 1)}.items():
    configure(pdl, side)
    pol.load_state_dict(state)
    pol.flat.m.zero_(); pol.flat.v.zero_()
    ag.step_count = 0
    ag.dev_step.zero_()
    step(0)
    torch.cuda.synchronize()
    g = pol.flat.g32.clone()
    st = ag.stats.clone()
    if ref is None:
        ref = (g, st)
    checks[name] = {"grad_max_abs_diff": float((g - ref[0]).abs().max()), "grad_max_abs": float(ref[0].abs().max()),
                    "stats_max_abs_diff": float((st - ref[1]).abs().max())}
print(json.dumps(checks, indent=1))

res = {}
order = [("base", 0, 0), ("pdl", 1, 0), ("side", 0, 1), ("pdl+side", 1, 1)]
for rep in range(3):
    for name, pdl, side in order:
        configure(pdl, side)
        for i in range(3):
            step(i)
        res.setdefault(name, []).append(timed(steps))
out = {"steps": steps, "ms_per_step": res, "checks": checks,
       "samples_per_s": {k: [round(B / (x * 1e-3)) for x in v] for k, v in res.items()}}
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ab_sched.json"), "w"), indent=1)
