"""per-launch breakdown of one PPO minibatch step (CUDA events around every launch)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_sc_b200 import ops, synthetic
from multimodal_sc_b200.model import SemComPolicy
from multimodal_sc_b200.ppo import PPOAgent
from multimodal_sc_b200.spec import SpecV0

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sp = SpecV0()
pol = SemComPolicy(sp)
ag = PPOAgent(pol, sp, minibatch=B)
m = synthetic.minibatch(sp, B, seed=1, device="cuda")
step = lambda: ag.minibatch_step(m["cam"], m["bev"], m["actions"], m["logp_old"], m["adv"], None, m["ret"], sync_stats=False)
for _ in range(3):
    step()
torch.cuda.synchronize()
prof = ops.profile()
prof.detail = True
with prof:
    step()
agg = {}
for n, ms, fl, by in prof.records:
    d = agg.setdefault(n, [0, 0.0, 0.0, 0.0])
    d[0] += 1; d[1] += ms; d[2] += fl; d[3] += by
tot = sum(d[1] for d in agg.values())
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
out = []
for n, (c, ms, fl, by) in rows:
    out.append(f"{n:34s} n={c:3d} ms={ms:7.3f} ({100*ms/tot:4.1f}%) per={ms/c*1e3:7.1f}us  TF={fl/ms/1e9 if ms else 0:7.1f} GB/s={by/ms/1e6 if ms else 0:7.1f}")
out.append(f"total {tot:.3f} ms")
print("\n".join(out))
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/step_breakdown.txt", "w").write("\n".join(out) + "\n")
