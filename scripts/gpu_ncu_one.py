"""Launch one sc_gemm_bf16 shape a few times (target for ncu captures under gpurun).

    python scripts/gpu_ncu_one.py <shape> [nwg] [gemm_dbg bits] [cta2]
shapes: see scripts/gpu_gemm_bottleneck.py:make
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch

from gpu_gemm_bottleneck import make
from multimodal_sc_b200 import ops

name = sys.argv[1]
nwg = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dbg = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cta2 = int(sys.argv[4]) if len(sys.argv) > 4 else 0
spec = make(name)
ops.debug_set("nwg", nwg)
ops.debug_set("gemm_dbg", dbg)
ops.debug_set("cta2", cta2)
A, kw = spec["A"](), spec["kw"]()
out = torch.empty(spec["M"], spec["N"], device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    ops.gemm(A, spec["B"], b_mn=spec.get("b_mn", False), out=out, **kw)
torch.cuda.synchronize()
print("ok", name, nwg, dbg)
