"""tiny PPO minibatch step for compute-sanitizer --tool memcheck (B = 2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_sc_b200 import synthetic
from multimodal_sc_b200.model import SemComPolicy
from multimodal_sc_b200.ppo import PPOAgent
from multimodal_sc_b200.spec import SpecV0
sp = SpecV0()
pol = SemComPolicy(sp)
ag = PPOAgent(pol, sp, minibatch=3, channel="rayleigh")
m = synthetic.minibatch(sp, 3, seed=1, device="cuda")
print(ag.minibatch_step(m["cam"], m["bev"], m["actions"], m["logp_old"], m["adv"], None, m["ret"]))
ro = synthetic.rollout(sp, 4, 2, seed=2)
print(PPOAgent(pol, sp, minibatch=4).update(ro))
