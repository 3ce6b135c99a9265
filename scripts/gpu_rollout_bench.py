"""act()/rollout throughput: eager launches vs the CUDA-graph replay, per number of environments; and the PPO step at
small minibatches, eager vs graph.  Writes gpurun_out/rollout_bench.json."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_sc_b200 import synthetic
from multimodal_sc_b200.model import SemComPolicy
from multimodal_sc_b200.ppo import PPOAgent
from multimodal_sc_b200.rollout import GraphedAct, RolloutCollector, SyntheticDrivingEnv
from multimodal_sc_b200.spec import SpecV0

sp = SpecV0()
out = {"act": {}, "ppo_step": {}}
pol = SemComPolicy(sp, seed=0)
for n in (32, 256, 1024):
    rec = {}
    for use_graph in (False, True):
        env = SyntheticDrivingEnv(n, device="cuda", seed=1)
        act = GraphedAct(pol, n, seed=3, use_graph=use_graph)
        cam, bev = env.observe()
        u = torch.rand(n, device="cuda")
        for _ in range(5):
            act(cam, bev, u)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        iters = 50
        for _ in range(iters):
            act(cam, bev, u)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / iters
        rec["graph" if use_graph else "eager"] = {"ms_per_act": round(dt * 1e3, 3), "env_steps_per_s": round(n / dt)}
    col = RolloutCollector(GraphedAct(pol, n, seed=3), SyntheticDrivingEnv(n, device="cuda", seed=1), horizon=16, device="cuda")
    col.collect()
    rec["collect_env_steps_per_s"] = round(RolloutCollector.steps_per_second(col))
    out["act"][str(n)] = rec
    print("act", n, json.dumps(rec), flush=True)
for B in (64, 256, 1024):
    rec = {}
    for graph in (False, True):
        p2 = SemComPolicy(sp, seed=0)
        ag = PPOAgent(p2, sp, minibatch=B, seed=5)
        m = synthetic.minibatch(sp, B, seed=1, device="cuda")
        step = lambda: ag.minibatch_step(m["cam"], m["bev"], m["actions"], m["logp_old"], m["adv"], None, m["ret"],
                                         sync_stats=False, graph=graph)
        for _ in range(4):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        iters = 20
        for _ in range(iters):
            step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / iters
        rec["graph" if graph else "eager"] = {"ms_per_step": round(dt * 1e3, 3), "samples_per_s": round(B / dt)}
    out["ppo_step"][str(B)] = rec
    print("ppo_step", B, json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/rollout_bench.json", "w"), indent=1)
