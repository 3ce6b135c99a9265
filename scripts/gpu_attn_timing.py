"""attention fwd/bwd timing at the workload's shapes, both forward variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_sc_b200 import ops

def t(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for B, S in ((4096, 64), (4096, 128)):
    H = 4
    qkv = torch.randn(B * S, 768, device="cuda").bfloat16()
    dout = torch.randn(B * S, 256, device="cuda").bfloat16()
    out, lse = ops.attn_fwd(qkv, B, S, H)
    dqkv = torch.empty_like(qkv)
    res = {}
    for st in (1, 2):
        ops.debug_set("attn_fwd_stages", st)
        res[f"fwd_stages{st}"] = t(lambda: ops.attn_fwd(qkv, B, S, H, out=out, lse=lse))
    ops.debug_set("attn_fwd_stages", 1)
    res["bwd"] = t(lambda: ops.attn_bwd(qkv, out, dout, lse, B, S, H, dqkv=dqkv))
    fb = qkv.numel() * 2 * 4 / 3; bb = qkv.numel() * 2 * 7 / 3
    print(f"B={B} S={S}: " + ", ".join(f"{k}={v:.1f}us" for k, v in res.items()),
          f"| fwd GB/s (best) {fb / min(res['fwd_stages1'], res['fwd_stages2']) / 1e3:.0f}, bwd GB/s {bb / res['bwd'] / 1e3:.0f}")
