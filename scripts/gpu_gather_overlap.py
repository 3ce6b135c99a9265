"""does the PCIe row gather (sc_gather_rows from pinned host memory, side stream) slow the PPO step it runs beside?
step alone, gather alone, both together, for several gather grid sizes -> gpurun_out/gather_overlap.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_sc_b200 import ops, synthetic
from multimodal_sc_b200.model import SemComPolicy
from multimodal_sc_b200.ppo import PPOAgent
from multimodal_sc_b200.spec import SpecV0

sp = SpecV0()
B = 4096
pol = SemComPolicy(sp)
ag = PPOAgent(pol, sp, minibatch=B)
mb = {k: v.cuda() for k, v in synthetic.minibatch(sp, B, seed=1, obs_dtype=torch.bfloat16).items()}
args = (mb["cam"], mb["bev"], mb["actions"], mb["logp_old"], mb["adv"], None, mb["ret"])
src = torch.empty((2 * B, 3, 128, 128), dtype=torch.bfloat16).pin_memory()
src.view(torch.int16).random_(0, 1000)
idx = torch.randperm(2 * B)[:B].cuda()
out = torch.empty((B, 3, 128, 128), dtype=torch.bfloat16, device="cuda")
side = torch.cuda.Stream()
res = {}
step = lambda: ag.minibatch_step(*args, sync_stats=False, graph="inplace")
for _ in range(5): step()
torch.cuda.synchronize()

def timed(with_step, ctas, reps=5):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    ev[0].record()
    if ctas:
        fork = torch.cuda.Event(); fork.record(); side.wait_event(fork)
        with torch.cuda.stream(side):
            ev[2].record(side)
            for _ in range(reps): ops.gather_rows(src, idx, out, max_ctas=ctas)
            ev[3].record(side)
    if with_step:
        for _ in range(reps): step()
    ev[1].record()
    torch.cuda.synchronize()
    return {"step_ms": ev[0].elapsed_time(ev[1]) / reps if with_step else None,
            "gather_ms": ev[2].elapsed_time(ev[3]) / reps if ctas else None}

res["step_alone"] = timed(True, 0)
for ctas in (2, 4, 8, 16):
    res[f"gather_alone_ctas{ctas}"] = timed(False, ctas)
    res[f"both_ctas{ctas}"] = timed(True, ctas)
res["step_alone_again"] = timed(True, 0)
print(json.dumps(res, indent=1))
json.dump(res, open("gpurun_out/gather_overlap.json", "w"), indent=1)
