"""Launch sc_mlp_bwd_dgrad (or sc_mlp_fwd with argv[1] == fwd) a few times on the step's shape: target for ncu captures."""
import os
import sys

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
mode = sys.argv[1] if len(sys.argv) > 1 else "bwd"
for _ in range(4):
    if mode == "fwd":
        ops.mlp_fwd(x, w1, b1, w2, b2, resid=resid, out=o)
    else:
        ops.mlp_bwd_dgrad(dy, w1, w2, hpre, dhpre=dhpre, dx=dx)
torch.cuda.synchronize()
print("ok", mode)
