"""Launch the fc2 weight-gradient GEMM (dW[256,1024] += dy^T h, split-K over 262144 tokens) a few times: target for
ncu captures.  argv[1] = 0 runs it as single CTAs (128x256 tiles), default = CTA pairs (256x256 per pair)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_sc_b200 import ops

pair = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ops.debug_set("cta2_wgrad", pair)
M, n_out, k_in = 262144, 256, 1024
g = lambda *s: (torch.randn(*s, device="cuda") * 0.1).to(torch.bfloat16)
dy, h = g(M, n_out), g(M, k_in)
dW = torch.zeros(n_out, k_in, device="cuda")
tiles = ((n_out + 127) // 128) * ((k_in + 255) // 256)
ks = max(1, 148 // tiles)
for _ in range(4):
    ops.gemm(dy, h, a_mn=True, b_mn=True, out_f32=dW, atomic=True, ksplit=ks)
torch.cuda.synchronize()
print("ok pair =", pair)
