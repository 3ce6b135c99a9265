"""Per-role wait accounting of the fused MLP kernels (needs a -DSC_MLP_ACCT build: SC_MLP_ACCT=1 python -m
multimodal_sc_b200.build --force).  CTA 0 prints each role's total cycles and the cycles inside its barrier waits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_sc_b200 import ops

BF16 = torch.bfloat16
D, R, M = 256, 1024, 262144
g = lambda *s: (torch.randn(*s, device="cuda") * 0.1).to(BF16)
x, resid, dy = g(M, D), g(M, D), g(M, D)
w1, w2 = g(R, D), g(D, R)
b1, b2 = torch.zeros(R, device="cuda"), torch.zeros(D, device="cuda")
hpre, dhpre = g(M, R), torch.empty(M, R, device="cuda", dtype=BF16)
o, dx = torch.empty(M, D, device="cuda", dtype=BF16), torch.empty(M, D, device="cuda", dtype=BF16)
for mode in (sys.argv[1:] or ["bwd", "fwd"]):
    run = (lambda: ops.mlp_bwd_dgrad(dy, w1, w2, hpre, dhpre=dhpre, dx=dx)) if mode == "bwd" else \
          (lambda: ops.mlp_fwd(x, w1, b1, w2, b2, resid=resid, out=o))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"=== mlp {mode}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us per launch (M {M}; {M // 128 / 148:.1f} tiles per CTA)", flush=True)
    ops.debug_set("mlp_dbg", 8)
    run()
    torch.cuda.synchronize()
    ops.debug_set("mlp_dbg", 0)
