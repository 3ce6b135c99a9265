"""cfg4 evaluation sweep on the GPU: python scripts/eval_snr_sweep.py [batch] -> gpurun_out/snr_sweep.json
(random-init weights and synthetic observations: the numbers exercise the tooling, they are not a trained result)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_sc_b200 import synthetic
from multimodal_sc_b200.evaluate import snr_sweep
from multimodal_sc_b200.model import SemComPolicy
from multimodal_sc_b200.spec import SpecV0

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sp = SpecV0()
pol = SemComPolicy(sp, init_std=0.05)
mb = synthetic.minibatch(sp, B, seed=6, device="cuda", obs_dtype=torch.bfloat16)
res = snr_sweep(pol, mb["cam"], mb["bev"], channels=("awgn", "rayleigh", "rician"))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/snr_sweep.json", "w"), indent=1)
for ch, rows in res.items():
    print(ch)
    for r in rows:
        print("  " + "  ".join(f"{k}={v:.4g}" for k, v in r.items()))
