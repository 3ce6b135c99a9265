"""one GEMM shape, a few launches — the target for `ncu --set full -k regex:gemm_tcgen05`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_sc_b200 import ops

shape = sys.argv[1] if len(sys.argv) > 1 else "fc1"
M = 262144
cfg = {"fc1": (1024, 256, dict(act="gelu", aux=True)), "proj": (256, 256, dict(resid=True)),
       "fc2": (256, 1024, dict(resid=True)), "qkv": (768, 256, {}), "dgelu": (1024, 256, dict(dgelu=True))}[shape]
N, K, kw = cfg
A = torch.randn(M, K, device="cuda").bfloat16()
B = torch.randn(N, K, device="cuda").bfloat16()
b_mn = False
if kw.get("dgelu"):
    B = torch.randn(K, N, device="cuda").bfloat16()     # dgrad form: W stored [K_red, N_out]
    b_mn = True
bias = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
args = dict(bias=bias, out=out, b_mn=b_mn)
if kw.get("dgelu"):
    args = dict(out=out, b_mn=True, aux_in=torch.randn(M, N, device="cuda").bfloat16())
if kw.get("aux"):
    args.update(act="gelu", aux_out=torch.empty_like(out))
if kw.get("resid"):
    args.update(resid=torch.randn(M, N, device="cuda").bfloat16())
for _ in range(4):
    ops.gemm(A, B, **args)
torch.cuda.synchronize()
print("ok", shape)
