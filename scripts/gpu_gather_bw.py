"""bandwidth of sc_gather_rows from pinned host memory (PCIe) and from device memory, against cudaMemcpyAsync of the same bytes"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_sc_b200 import ops

n_src, n = 8192, 4096
src = torch.empty((n_src, 3, 128, 128), dtype=torch.bfloat16).pin_memory()
src.view(torch.int16).random_(0, 1000)
idx = torch.randperm(n_src)[:n].cuda()
out = torch.empty((n, 3, 128, 128), dtype=torch.bfloat16, device="cuda")
nbytes = out.numel() * 2
res = {}

def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

ms = t(lambda: out.copy_(src[:n], non_blocking=True))
res["cudaMemcpyAsync_pinned_contiguous"] = {"ms": ms, "GBps": nbytes / ms / 1e6}
for ctas in (148, 296, 592, 1184, 2368):
    ms = t(lambda: ops.gather_rows(src, idx, out, max_ctas=ctas))
    res[f"gather_pinned_ctas{ctas}"] = {"ms": ms, "GBps": nbytes / ms / 1e6}
ms = t(lambda: ops.gather_rows(src, None, out[:n], max_ctas=592))
res["copy_pinned_noidx_ctas592"] = {"ms": ms, "GBps": nbytes / ms / 1e6}
dsrc = src.cuda()
for ctas in (148, 1184):
    ms = t(lambda: ops.gather_rows(dsrc, idx, out, max_ctas=ctas))
    res[f"gather_device_ctas{ctas}"] = {"ms": ms, "GBps": 2 * nbytes / ms / 1e6}
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gather_bw.json", "w"), indent=1)
