"""ncu target: sc_dp_reduce_adam over the model's flat bucket (5.17 M parameters) with world = 1 — the local HBM side of
the kernel (30 B per element: p, g, m, v read; p, m, v, bf16 shadow written).  The NVLink side cannot be profiled under
ncu (it serialises kernels, a multi-rank barrier would never complete); it is timed by scripts/gpu_dp_fused_check.py."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_sc_b200 import _lib
from multimodal_sc_b200.peer import preload_kernels

preload_kernels("cuda")
n = 5166016
p_, g_, m_, v_ = (torch.randn(n, device="cuda") for _ in range(4))
m_.abs_(); v_.abs_()
pb = torch.empty(n, device="cuda", dtype=torch.bfloat16)
arr = lambda *ts: (C.c_void_p * len(ts))(*[x.data_ptr() for x in ts])
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for i in range(4):
    flush.zero_()
    _lib.check(lib.sc_dp_reduce_adam(arr(g_), arr(p_), arr(pb), m_.data_ptr(), v_.data_ptr(), 1, 0, n, 3e-4, 0.9, 0.999,
                                     1e-8, 3 + i, None, 0, st))
torch.cuda.synchronize()
print("ok")
