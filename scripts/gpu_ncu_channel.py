"""Launch sc_pnorm_channel_fwd / bwd at the bench shape (4096 x 2 x 1024 fp32 symbols) a few times: ncu target."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_sc_b200 import ops

B, k = (int(sys.argv[2]) if len(sys.argv) > 2 else 4096), 1024
kind = sys.argv[1] if len(sys.argv) > 1 else "awgn"
z = torch.randn(2, B, k, device="cuda")
y16 = torch.empty(B, 2, k, device="cuda", dtype=torch.bfloat16)
scale = torch.empty(B, 2, device="cuda")
dy = torch.randn(B, 2, k, device="cuda")
dz = torch.empty(2, B, k, device="cuda", dtype=torch.bfloat16)
for i in range(4):
    ops.pnorm_channel_fwd(z, kind, 10.0, seed=1234 + i, sample_offset=0, noise=None, fading=None, eps=1e-12, y_f32=None,
                          y_bf16=y16, scale=scale, z_strides=(k, B * k, B, k))
    ops.pnorm_channel_bwd(dy, z, scale, kind, 10.0, seed=1234 + i, sample_offset=0, fading=None, dz_bf16=dz,
                          z_strides=(k, B * k, B, k))
torch.cuda.synchronize()
print("ok", kind)
