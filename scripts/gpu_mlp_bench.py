"""Fused MLP kernels vs the two-GEMM path on the step's shapes (run under gpurun); writes gpurun_out/mlp_bench.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_sc_b200 import ops

BF16 = torch.bfloat16
D, R = 256, 1024


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    out = {}
    for M in (262144, 524288):
        g = lambda *s: (torch.randn(*s, device="cuda") * 0.1).to(BF16)
        x, resid, dy = g(M, D), g(M, D), g(M, D)
        w1, w2 = g(R, D), g(D, R)
        b1, b2 = torch.zeros(R, device="cuda"), torch.zeros(D, device="cuda")
        h, hpre, dhpre = torch.empty(M, R, device="cuda", dtype=BF16), torch.empty(M, R, device="cuda", dtype=BF16), torch.empty(M, R, device="cuda", dtype=BF16)
        o, dx = torch.empty(M, D, device="cuda", dtype=BF16), torch.empty(M, D, device="cuda", dtype=BF16)

        def unfused_fwd():
            ops.gemm(x, w1, bias=b1, act="gelu", aux_out=hpre, out=h)
            ops.gemm(h, w2, bias=b2, resid=resid, out=o)

        def fused_fwd():
            ops.mlp_fwd(x, w1, b1, w2, b2, resid=resid, hpre=hpre, h=h, out=o)

        def fused_fwd_nosave():
            ops.mlp_fwd(x, w1, b1, w2, b2, resid=resid, out=o)

        def fused_fwd_h_only():
            ops.mlp_fwd(x, w1, b1, w2, b2, resid=resid, h=h, out=o)

        def fused_fwd_hpre_only():
            ops.mlp_fwd(x, w1, b1, w2, b2, resid=resid, hpre=hpre, out=o)

        def unfused_bwd():
            ops.gemm(dy, w2, b_mn=True, out=dhpre, aux_in=hpre)
            ops.gemm(dhpre, w1, b_mn=True, out=dx)

        def fused_bwd():
            ops.mlp_bwd_dgrad(dy, w1, w2, hpre, dhpre=dhpre, dx=dx)

        unfused_fwd()
        rec = {}
        for name, fn in (("unfused_fwd", unfused_fwd), ("fused_fwd", fused_fwd), ("fused_fwd_nosave", fused_fwd_nosave), ("fused_fwd_h_only", fused_fwd_h_only), ("fused_fwd_hpre_only", fused_fwd_hpre_only),
                         ("unfused_bwd_dgrad", unfused_bwd), ("fused_bwd_dgrad", fused_bwd)):
            us = timeit(fn)
            rec[name] = {"us": round(us, 1), "tflops": round(4.0 * M * D * R / us / 1e6, 1)}
        out[str(M)] = rec
        print(M, json.dumps(rec), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/mlp_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
