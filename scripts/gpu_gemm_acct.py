"""Per-role wait accounting of the GEMM pipeline (gemm_dbg bit 512): for the step's main shapes CTA 0 prints how many
cycles the producer, the MMA warp and the epilogue warpgroups spent in total and inside their barrier waits."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

from gpu_gemm_bottleneck import make
from multimodal_sc_b200 import ops

names = sys.argv[1:] or ["qkv", "fc1", "proj", "proj_ln", "fc2_ln", "dgrad_fc1"]
for name in names:
    try:
        spec = make(name)
    except Exception as e:  # noqa: BLE001
        print("skip", name, e)
        continue
    A, kw = spec["A"](), spec["kw"]()
    out = torch.empty(spec["M"], spec["N"], device="cuda", dtype=torch.bfloat16)
    run = lambda: ops.gemm(A, spec["B"], b_mn=spec.get("b_mn", False), out=out, **kw)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"=== {name}: M {spec['M']} N {spec['N']} K {A.shape[1]}  {e0.elapsed_time(e1) / 5 * 1e3:.1f} us per launch", flush=True)
    ops.debug_set("gemm_dbg", 512)
    run()
    torch.cuda.synchronize()
    ops.debug_set("gemm_dbg", 0)
