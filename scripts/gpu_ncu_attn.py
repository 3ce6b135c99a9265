"""attention forward + backward at the encoder shape (4096 samples x 64 tokens, 4 heads): ncu target."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_sc_b200 import ops

B, S, H = 4096, 64, 4
qkv = torch.randn(B * S, 768, device="cuda").bfloat16()
dout = torch.randn(B * S, 256, device="cuda").bfloat16()
out, lse = ops.attn_fwd(qkv, B, S, H)
dqkv = torch.empty_like(qkv)
for _ in range(3):
    ops.attn_fwd(qkv, B, S, H, out=out, lse=lse)
    ops.attn_bwd(qkv, out, dout, lse, B, S, H, dqkv=dqkv)
torch.cuda.synchronize()
print("ok")
