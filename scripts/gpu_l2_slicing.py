"""Does running a transformer block slice by slice (working set inside the 126 MB L2) beat the full-batch schedule?

Captures one encoder-block forward (LN1, qkv, attention, proj+resid, LN2, fc1+GELU, fc2+resid; training saves) over
M = 262144 token rows as a CUDA graph, once as 7 whole-batch launches and once per row slice, and times the replays.
Writes gpurun_out/l2_slicing.json.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_sc_b200 import ops

BF16, F32 = torch.bfloat16, torch.float32
D, R, H, S = 256, 1024, 4, 64
B = 4096
M = B * S


def main():
    g = lambda *s, sc=0.1: (torch.randn(*s, device="cuda") * sc).to(BF16)
    x = g(M, D, sc=1.0)
    W = dict(qkv=g(3 * D, D), proj=g(D, D), fc1=g(R, D), fc2=g(D, R))
    bias = dict(qkv=torch.zeros(3 * D, device="cuda"), proj=torch.zeros(D, device="cuda"), fc1=torch.zeros(R, device="cuda"),
                fc2=torch.zeros(D, device="cuda"))
    gam, bet = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    e = lambda n, dt=BF16: torch.empty(M, n, device="cuda", dtype=dt)
    buf = dict(xn1=e(D), qkv=e(3 * D), ao=e(D), x1=e(D), xn2=e(D), h=e(R), hpre=e(R), out=e(D),
               mean1=torch.empty(M, device="cuda"), rstd1=torch.empty(M, device="cuda"),
               mean2=torch.empty(M, device="cuda"), rstd2=torch.empty(M, device="cuda"), lse=torch.empty(M, H, device="cuda"))

    def block(r0, r1):
        sl = slice(r0, r1)
        nb = (r1 - r0) // S
        ops.layernorm_fwd(x[sl], gam, bet, 1e-5, y=buf["xn1"][sl], mean=buf["mean1"][sl], rstd=buf["rstd1"][sl])
        ops.gemm(buf["xn1"][sl], W["qkv"], bias=bias["qkv"], out=buf["qkv"][sl])
        ops.attn_fwd(buf["qkv"][sl], nb, S, H, out=buf["ao"][sl], lse=buf["lse"][sl])
        ops.gemm(buf["ao"][sl], W["proj"], bias=bias["proj"], resid=x[sl], out=buf["x1"][sl])
        ops.layernorm_fwd(buf["x1"][sl], gam, bet, 1e-5, y=buf["xn2"][sl], mean=buf["mean2"][sl], rstd=buf["rstd2"][sl])
        ops.gemm(buf["xn2"][sl], W["fc1"], bias=bias["fc1"], act="gelu", aux_out=buf["hpre"][sl], out=buf["h"][sl])
        ops.gemm(buf["h"][sl], W["fc2"], bias=bias["fc2"], resid=buf["x1"][sl], out=buf["out"][sl])

    def run(rows):
        for r0 in range(0, M, rows):
            block(r0, min(M, r0 + rows))

    res = {}
    for rows in (M, 148 * 128 * 4, 148 * 128 * 2, 148 * 128):
        run(rows)                                    # warm-up (sets kernel attributes outside the capture)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            run(rows)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        res[str(rows)] = round(e0.elapsed_time(e1) / 10 * 1e3, 1)
        print("rows per slice", rows, "slices", -(-M // rows), "block fwd us", res[str(rows)], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/l2_slicing.json", "w"), indent=1)


if __name__ == "__main__":
    main()
