"""Bottleneck isolation for sc_gemm_bf16 on the step's dominant shapes (run under gpurun).

For every shape x variant the kernel is timed (CUDA events, inputs larger than L2 by rotating over
several operand sets) with parts of the pipeline switched off through sc_debug_set("gemm_dbg", bits):
  1 = epilogue only releases the accumulator     -> loads + MMA alone
  2 = epilogue math + smem staging, no TMA stores/loads -> cost of the global writes / epilogue inputs
  4 = no MMAs                                    -> loads + epilogue alone
  8 = A always block 0 (L2 hits)                 -> cost of the HBM reads
 16 = B not loaded                               -> cost of re-streaming the weights
Results with any bit set are numerically wrong by construction; this is a timing probe only.
Writes gpurun_out/gemm_bottleneck.json.
"""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from multimodal_sc_b200 import ops

BF16 = torch.bfloat16
M = 262144
SETS = 3          # operand sets rotated between launches (each set >> 126 MB L2 for the big shapes)


def make(shape_name):
    g = lambda *s: (torch.randn(*s, device="cuda") * 0.1).to(BF16)
    if shape_name == "qkv":      # [M,256] x [768,256]^T + bias
        return dict(M=M, N=768, K=256, A=lambda: g(M, 256), B=g(768, 256), kw=lambda: dict(bias=torch.zeros(768, device="cuda")))
    if shape_name == "proj":     # + bias + residual
        return dict(M=M, N=256, K=256, A=lambda: g(M, 256), B=g(256, 256),
                    kw=lambda: dict(bias=torch.zeros(256, device="cuda"), resid=g(M, 256)))
    if shape_name == "fc1":      # + bias, GELU, pre-activation copy
        return dict(M=M, N=1024, K=256, A=lambda: g(M, 256), B=g(1024, 256),
                    kw=lambda: dict(bias=torch.zeros(1024, device="cuda"), act="gelu", aux_out=torch.empty(M, 1024, device="cuda", dtype=BF16)))
    if shape_name == "fc2":      # K = 1024, + bias + residual
        return dict(M=M, N=256, K=1024, A=lambda: g(M, 1024), B=g(256, 1024),
                    kw=lambda: dict(bias=torch.zeros(256, device="cuda"), resid=g(M, 256)))
    if shape_name == "dgrad_fc2":  # dh = dy W2, * GELU'(hpre)
        return dict(M=M, N=1024, K=256, A=lambda: g(M, 256), B=g(256, 1024), b_mn=True,
                    kw=lambda: dict(aux_in=g(M, 1024)))
    if shape_name == "dgrad_fc1":  # dx = dhpre W1
        return dict(M=M, N=256, K=1024, A=lambda: g(M, 1024), B=g(1024, 256), b_mn=True, kw=lambda: dict())
    if shape_name == "dgrad_proj":  # dao = dx1 W_proj: K = 256, plain bf16 output
        return dict(M=M, N=256, K=256, A=lambda: g(M, 256), B=g(256, 256), b_mn=True, kw=lambda: dict())
    if shape_name == "dgrad_qkv":  # dxn1 = dqkv W_qkv: K = 768
        return dict(M=M, N=256, K=768, A=lambda: g(M, 768), B=g(768, 256), b_mn=True, kw=lambda: dict())
    if shape_name == "sym":        # symbol projection: [M,256] x [16,256]^T + bias -> fp32 [M,16]
        return dict(M=M, N=16, K=256, A=lambda: g(M, 256), B=g(16, 256), f32=True,
                    kw=lambda: dict(bias=torch.zeros(16, device="cuda")))
    if shape_name == "dsym":       # its dgrad twin in the decoder: [2M,256] x [256,16] -> fp32 [2M,16]
        return dict(M=2 * M, N=16, K=256, A=lambda: g(2 * M, 256), B=g(256, 16), b_mn=True, f32=True, kw=lambda: dict())
    if shape_name == "inproj":     # decoder input projection: K = 16
        return dict(M=2 * M, N=256, K=16, A=lambda: g(2 * M, 16), B=g(256, 16),
                    kw=lambda: dict(bias=torch.zeros(256, device="cuda")))
    if shape_name in ("wgrad_fc1", "wgrad_fc2", "wgrad_qkv"):   # dW[n_out, k_in] += dy^T x, split-K over the tokens, fp32 atomics
        n_out, k_in = {"wgrad_fc1": (1024, 256), "wgrad_fc2": (256, 1024), "wgrad_qkv": (768, 256)}[shape_name]
        return dict(M=n_out, N=k_in, K=M, A=lambda: g(M, n_out), B=g(M, k_in), wgrad=True, kw=lambda: dict())
    if shape_name in ("proj_ln", "fc2_ln"):   # residual GEMM with the following LayerNorm fused into the epilogue
        K = 256 if shape_name == "proj_ln" else 1024
        def kw():
            return dict(bias=torch.zeros(256, device="cuda"), resid=g(M, 256),
                        ln=dict(gamma=torch.ones(256, device="cuda"), beta=torch.zeros(256, device="cuda"), eps=1e-5,
                                y=torch.empty(M, 256, device="cuda", dtype=BF16), mean=torch.empty(M, device="cuda"),
                                rstd=torch.empty(M, device="cuda")))
        return dict(M=M, N=256, K=K, A=lambda: g(M, K), B=g(256, K), kw=kw)
    if shape_name == "big":
        return dict(M=8192, N=8192, K=8192, A=lambda: g(8192, 8192), B=g(8192, 8192), kw=lambda: dict())
    if shape_name == "plain256":
        return dict(M=M, N=256, K=256, A=lambda: g(M, 256), B=g(256, 256), kw=lambda: dict())
    raise KeyError(shape_name)


def time_case(spec, dbg, nwg, iters=12, slots=0, cta2=0):
    ops.debug_set("reset", 0)
    ops.debug_set("cta2", cta2)
    ops.debug_set("gemm_dbg", dbg)
    ops.debug_set("nwg", nwg)
    ops.debug_set("slots", slots)
    sets = []
    wg_out = torch.zeros(spec["M"], spec["N"], device="cuda") if spec.get("wgrad") else None
    for _ in range(SETS):
        kw = spec["kw"]()
        sets.append((spec["A"](), kw, None if spec.get("wgrad") else
                     torch.empty(spec["M"], spec["N"], device="cuda", dtype=torch.float32 if spec.get("f32") else BF16)))
    def run(i):
        A, kw, out = sets[i % SETS]
        if spec.get("wgrad"):
            tiles = ((spec["M"] + 127) // 128) * ((spec["N"] + 255) // 256)
            ops.gemm(A, spec["B"], a_mn=True, b_mn=True, out_f32=wg_out, atomic=True, ksplit=max(1, 148 // tiles))
        elif spec.get("f32"):
            ops.gemm(A, spec["B"], b_mn=spec.get("b_mn", False), out_f32=out, **kw)
        else:
            ops.gemm(A, spec["B"], b_mn=spec.get("b_mn", False), out=out, **kw)
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    ops.debug_set("reset", 0)
    return e0.elapsed_time(e1) / iters * 1e3   # us


def main():
    out = {}
    variants = [("full", 0), ("epi_skip", 1), ("epi_no_tma", 2), ("no_mma", 4), ("a_l2", 8), ("no_b", 16),
                ("epi_skip+a_l2", 9), ("epi_skip+no_b", 17), ("epi_skip+no_mma", 5), ("no_mma+no_tma", 6)]
    if os.environ.get("SC_BOTTLENECK_QUICK"):
        variants = [("full", 0), ("epi_skip", 1), ("epi_no_tma", 2), ("no_mma+no_tma", 6)]
    names = ("qkv", "proj", "fc1", "fc2", "dgrad_fc2", "dgrad_fc1", "plain256")
    if os.environ.get("SC_BOTTLENECK_MMA"):
        names = ("fc2", "qkv", "fc1", "big")
    if os.environ.get("SC_BOTTLENECK_SKINNY"):
        names = ("sym", "dsym", "inproj", "plain256")
    if os.environ.get("SC_BOTTLENECK_WGRAD"):
        names = ("wgrad_fc1", "wgrad_fc2", "wgrad_qkv", "fc2")
    if os.environ.get("SC_BOTTLENECK_LN"):
        names = ("proj", "proj_ln", "fc2", "fc2_ln")
    for name in names:
        spec = make(name)
        rec = {}
        if os.environ.get("SC_BOTTLENECK_MMA"):
            # raw tensor-pipe rate: no loads, no epilogue work (bits 32 | 1), single CTA per tile and CTA pairs
            for cta2 in (0, 1):
                for vn, bits in (("mma_only", 33), ("loads+mma", 1), ("full", 0), ("spin.mma_only", 97), ("spin.loads+mma", 65), ("spin.full", 64)):
                    ops.debug_set("reset", 0)
                    try:
                        rec[f"cta2={cta2}.{vn}"] = round(time_case(spec, bits, 1, slots=2, cta2=cta2), 1)
                    except Exception as e:   # noqa: BLE001
                        rec[f"cta2={cta2}.{vn}"] = f"ERR {e}"
            fl = 2.0 * spec["M"] * spec["N"] * spec["K"]
            rec["tflops_mma_only"] = round(fl / rec["cta2=0.mma_only"] / 1e6, 1) if isinstance(rec["cta2=0.mma_only"], float) else None
            out[name] = rec
            print(name, json.dumps(rec), flush=True)
            continue
        if os.environ.get("SC_BOTTLENECK_SKINNY") or os.environ.get("SC_BOTTLENECK_WGRAD") or os.environ.get("SC_BOTTLENECK_LN"):
            for vn, bits in (("full", 0), ("epi_skip", 1), ("no_mma", 4), ("no_b", 16), ("epi_skip+no_b", 17), ("epi_skip+no_mma", 5),
                             ("loads_off", 32), ("loads_off+epi_skip", 33)):
                try:
                    rec[vn] = round(time_case(spec, bits, 0), 1)
                except Exception as e:   # noqa: BLE001
                    rec[vn] = f"ERR {e}"
            out[name] = rec
            print(name, json.dumps(rec), flush=True)
            continue
        if os.environ.get("SC_BOTTLENECK_CFG"):
            # epilogue configuration sweep: (warpgroups, slots per warp); 0/0 = the library's own choice
            for nwg, slots in ((0, 0), (1, 2), (2, 2), (1, 4), (2, 4)):
                for vn, bits in (("full", 0), ("epi_skip", 1), ("no_mma+no_tma", 6)):
                    try:
                        rec[f"w{nwg}s{slots}.{vn}"] = round(time_case(spec, bits, nwg, slots=slots), 1)
                    except Exception as e:   # noqa: BLE001
                        rec[f"w{nwg}s{slots}.{vn}"] = f"ERR {e}"
            out[name] = rec
            print(name, json.dumps(rec), flush=True)
            continue
        for nwg in (1, 2):
            for vn, bits in variants:
                try:
                    rec[f"nwg{nwg}.{vn}"] = round(time_case(spec, bits, nwg), 1)
                except Exception as e:   # noqa: BLE001
                    rec[f"nwg{nwg}.{vn}"] = f"ERR {e}"
        out[name] = rec
        print(name, json.dumps(rec), flush=True)
        del spec
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gemm_bottleneck.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
