"""cfg2 forward (B = 256, AWGN 10 dB, inference schedule) a few times — the command profiled for the small-batch launch list
(ncu --metrics gpu__time_duration.sum)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_sc_b200 import synthetic
from multimodal_sc_b200.model import SemComPolicy
from multimodal_sc_b200.spec import SpecV0

sp = SpecV0()
pol = SemComPolicy(sp)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mb = synthetic.minibatch(sp, B, seed=5, device="cuda", obs_dtype=torch.bfloat16)
for _ in range(3):
    pol.engine.forward(mb["cam"], mb["bev"], 10.0, "awgn", 77, None, None, 0, train=False)
torch.cuda.synchronize()
