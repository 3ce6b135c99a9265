"""Bring-up probe for sc_gemm_bf16 on a real B200 (run under gpurun).

    python scripts/gpu_probe_gemm.py --case <name>     # one case, exit code 0/1
    python scripts/gpu_probe_gemm.py --all             # every case, each in its own subprocess

Each case appends one JSON line to gpurun_out/gemm_probe.jsonl; a trap in one case cannot
poison the next because every case gets a fresh CUDA context.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

OUT = "gpurun_out/gemm_probe.jsonl"


def ref_gemm(A, B, a_mn, b_mn):
    a = A.float().t() if a_mn else A.float()
    b = B.float().t() if b_mn else B.float()
    return a @ b.t()


def gelu_tanh(x):
    import torch
    return torch.nn.functional.gelu(x, approximate="tanh")


def dgelu_tanh(x):
    import torch
    x = x.detach().clone().requires_grad_(True)
    y = gelu_tanh(x)
    (g,) = torch.autograd.grad(y.sum(), x)
    return g


CASES = {}


def case(name):
    def deco(fn):
        CASES[name] = fn
        return fn
    return deco


def _mk(M, N, K, a_mn, b_mn, seed=0, scale=1.0):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = (torch.randn((K, M) if a_mn else (M, K), generator=g, device="cuda") * scale).bfloat16()
    B = (torch.randn((K, N) if b_mn else (N, K), generator=g, device="cuda") * scale).bfloat16()
    return A, B


def _report(name, got, ref, extra=None, tol=2e-2):
    import torch
    got = got.float()
    err = (got - ref).abs()
    denom = ref.abs().max().clamp_min(1e-6)
    rel = float(err.max() / denom)
    bad = err > tol * denom
    rec = {"case": name, "max_rel_err": rel, "frac_bad": float(bad.float().mean()), "ok": rel < tol,
           "has_nan": bool(torch.isnan(got).any())}
    if not rec["ok"]:
        # localise: bad fraction by row%8, by row//32 (first 8), by col//64 (first 8)
        M, N = ref.shape
        r = torch.arange(M, device=ref.device)
        c = torch.arange(N, device=ref.device)
        rec["bad_by_row_mod8"] = [float(bad[r % 8 == i].float().mean()) for i in range(8)]
        rec["bad_by_row_div32"] = [float(bad[(r // 32) == i].float().mean()) for i in range(min(8, (M + 31) // 32))]
        rec["bad_by_col_div64"] = [float(bad[:, (c // 64) == i].float().mean()) for i in range(min(8, (N + 63) // 64))]
        rec["bad_by_col_mod8"] = [float(bad[:, c % 8 == i].float().mean()) for i in range(8)]
        rec["sample_got"] = got[:2, :4].tolist()
        rec["sample_ref"] = ref[:2, :4].tolist()
    if extra:
        rec.update(extra)
    return rec


def _plain(name, M, N, K, a_mn, b_mn, ksplit=1, dbg=None):
    import torch
    from multimodal_sc_b200 import ops
    if dbg:
        for k, v in dbg.items():
            ops.debug_set(k, v)
    A, B = _mk(M, N, K, a_mn, b_mn)
    ref = ref_gemm(A, B, a_mn, b_mn)
    if ksplit > 1:
        out = torch.zeros(M, N, device="cuda")
        ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out_f32=out, atomic=True, ksplit=ksplit)
    else:
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out=out)
    torch.cuda.synchronize()
    return _report(name, out, ref, {"M": M, "N": N, "K": K, "a_mn": a_mn, "b_mn": b_mn, "ksplit": ksplit,
                                    "dbg": dbg})


# progressive K-major/K-major cases
for nm, (M, N, K) in {"kk_k16": (128, 256, 16), "kk_k64": (128, 256, 64), "kk_k256": (128, 256, 256),
                      "kk_k1024": (256, 512, 1024), "kk_big": (4096, 1024, 256),
                      "kk_ragged": (200, 72, 40), "kk_n16": (300, 16, 256), "kk_k16in": (384, 256, 16)}.items():
    CASES[nm] = (lambda nm=nm, M=M, N=N, K=K: _plain(nm, M, N, K, False, False))
# dgrad form (A K-major, B MN-major)
for nm, (M, N, K) in {"kmn_k16": (128, 256, 16), "kmn_k64": (128, 256, 64), "kmn_k256": (256, 1024, 256),
                      "kmn_ragged": (200, 64, 1024)}.items():
    CASES[nm] = (lambda nm=nm, M=M, N=N, K=K: _plain(nm, M, N, K, False, True))
# wgrad form (both MN-major), with split-K
for nm, (M, N, K, ks) in {"mnmn_k16": (128, 256, 16, 1), "mnmn_k64": (128, 256, 64, 1),
                          "mnmn_k4096": (256, 256, 4096, 1), "mnmn_split": (256, 1024, 8192, 16),
                          "mnmn_n64": (256, 64, 8192, 37), "mnmn_m16": (16, 256, 4096, 8)}.items():
    CASES[nm] = (lambda nm=nm, M=M, N=N, K=K, ks=ks: _plain(nm, M, N, K, True, True, ks))
CASES["mnk_k64"] = lambda: _plain("mnk_k64", 128, 256, 64, True, False)
# alternative descriptor hypotheses for MN-major (only consulted if the defaults fail)
CASES["alt_kmn_swap"] = lambda: _plain("alt_kmn_swap", 128, 256, 64, False, True, dbg={"b_lbo": 1024, "b_sbo": 8192})
CASES["alt_mnmn_swap"] = lambda: _plain("alt_mnmn_swap", 128, 256, 64, True, True,
                                        dbg={"a_lbo": 1024, "a_sbo": 8192, "b_lbo": 1024, "b_sbo": 8192})


@case("epi_bias_gelu_aux")
def _():
    import torch
    from multimodal_sc_b200 import ops
    M, N, K = 512, 1024, 256
    A, B = _mk(M, N, K, False, False, scale=0.3)
    bias = torch.randn(N, device="cuda")
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out = torch.empty_like(pre)
    ops.gemm(A, B, bias=bias, act="gelu", aux_out=pre, out=out)
    ref_pre = ref_gemm(A, B, False, False) + bias
    r1 = _report("epi_bias_gelu_aux.pre", pre, ref_pre)
    r2 = _report("epi_bias_gelu_aux", out, gelu_tanh(ref_pre))
    r2["ok"] = r1["ok"] and r2["ok"]
    r2["pre_rel"] = r1["max_rel_err"]
    return r2


@case("epi_resid_rowadd")
def _():
    import torch
    from multimodal_sc_b200 import ops
    M, N, K = 640, 256, 768
    A, B = _mk(M, N, K, False, False, scale=0.2)
    bias = torch.randn(N, device="cuda")
    pos = torch.randn(64, N, device="cuda")
    resid = torch.randn(M, N, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, bias=bias, rowadd=pos, resid=resid, out=out)
    ref = ref_gemm(A, B, False, False) + bias + pos.repeat(M // 64, 1) + resid.float()
    return _report("epi_resid_rowadd", out, ref)


@case("epi_dgelu_rowgrp_f32")
def _():
    import torch
    from multimodal_sc_b200 import ops
    M, N, K = 512, 1024, 256
    A, B = _mk(M, N, K, False, True, scale=0.3)
    pre = torch.randn(M, N, device="cuda").bfloat16()
    grp = torch.randn(M // 128, N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    o32 = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, b_mn=True, rowgrp=grp, rowgrp_size=128, rowgrp_scale=0.5, aux_in=pre, alpha=2.0, out=out,
             out_f32=o32)
    ref = (2.0 * ref_gemm(A, B, False, True) + 0.5 * grp.repeat_interleave(128, 0)) * dgelu_tanh(pre.float())
    r = _report("epi_dgelu_rowgrp_f32", out, ref)
    r2 = _report("f32", o32, ref)
    r["ok"] = r["ok"] and r2["ok"]
    r["f32_rel"] = r2["max_rel_err"]
    return r


@case("epi_tanh_scalar_n10")
def _():
    import torch
    from multimodal_sc_b200 import ops
    M, N, K = 333, 10, 256
    A, B = _mk(M, N, K, False, False, scale=0.2)
    bias = torch.randn(N, device="cuda")
    o32 = torch.zeros(M, 16, device="cuda")
    ops.gemm(A, B, bias=bias, act="tanh", out_f32=o32[:, :N])
    ref = torch.tanh(ref_gemm(A, B, False, False) + bias)
    r = _report("epi_tanh_scalar_n10", o32[:, :N], ref)
    r["pad_untouched"] = bool((o32[:, N:] == 0).all())
    r["ok"] = r["ok"] and r["pad_untouched"]
    return r


@case("staged_tail_rows")
def _():
    """staged epilogue with an M tail (TMA clips), residual input and row-group add"""
    import torch
    from multimodal_sc_b200 import ops
    M, N, K = 1000, 512, 256
    A, B = _mk(M, N, K, False, False, scale=0.3)
    bias = torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda").bfloat16()
    grp = torch.randn((M + 127) // 128, N, device="cuda")
    out = torch.full((M + 8, N), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, bias=bias, resid=resid, rowgrp=grp, rowgrp_size=128, rowgrp_scale=0.25, out=out[:M])
    ref = ref_gemm(A, B, False, False) + bias + resid.float() + 0.25 * grp.repeat_interleave(128, 0)[:M]
    r = _report("staged_tail_rows", out[:M], ref)
    r["guard_rows_untouched"] = bool((out[M:] == 7.0).all())
    r["ok"] = r["ok"] and r["guard_rows_untouched"]
    return r


@case("wgrad_colsum")
def _():
    """bias gradient fused into the wgrad GEMM (ones-tile MMA), several shapes incl. ragged rows"""
    import torch
    from multimodal_sc_b200 import ops
    oks, rels = [], {}
    for nm, Nl, Kl, T, ks in (("fc1", 1024, 256, 8192, 9), ("sym", 16, 256, 4096, 8), ("qkv", 768, 256, 5000, 5),
                              ("patchlidar", 256, 64, 4096, 3), ("nosplit", 256, 256, 512, 1)):
        g = torch.Generator(device="cuda").manual_seed(1)
        dy = torch.randn(T, Nl, generator=g, device="cuda").bfloat16()
        x = torch.randn(T, Kl, generator=g, device="cuda").bfloat16()
        dW = torch.zeros(Nl, Kl, device="cuda")
        db = torch.ones(Nl, device="cuda")
        ops.gemm(dy, x, a_mn=True, b_mn=True, out_f32=dW, atomic=True, ksplit=ks, colsum=db)
        r1 = _report(nm + ".dW", dW, dy.float().t() @ x.float())
        r2 = _report(nm + ".db", db, 1.0 + dy.float().sum(0), tol=2e-3)
        oks += [r1["ok"], r2["ok"]]
        rels[nm] = (r1["max_rel_err"], r2["max_rel_err"])
    return {"case": "wgrad_colsum", "ok": all(oks), "rels": rels}


@case("staged_pos_bf16")
def _():
    """pos-embed add through the TMA input-tile path (bf16 table, row % period)"""
    import torch
    from multimodal_sc_b200 import ops
    oks, rels = [], {}
    for nm, M, N, K, period in (("cam", 64 * 37, 256, 768, 64), ("dec", 128 * 21, 256, 16, 128)):
        A, B = _mk(M, N, K, False, False, scale=0.2)
        bias = torch.randn(N, device="cuda")
        pos = torch.randn(period, N, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ops.gemm(A, B, bias=bias, rowadd=pos, out=out)
        ref = ref_gemm(A, B, False, False) + bias + pos.float().repeat(M // period, 1)
        r = _report(nm, out, ref)
        oks.append(r["ok"]); rels[nm] = r["max_rel_err"]
    return {"case": "staged_pos_bf16", "ok": all(oks), "rels": rels}


@case("staged_many_tiles")
def _():
    """more tiles than CTAs so the slot ring / input prefetch wraps many times; all three staged modes"""
    import torch
    from multimodal_sc_b200 import ops
    M, K = 128 * 700, 256
    oks, rels = [], {}
    for nm, N, kw in (("out1", 768, {}), ("out2", 1024, {"act": "gelu", "aux": True}), ("in_resid", 256, {"resid": True}),
                      ("in_dgelu", 1024, {"aux_in": True})):
        A, B = _mk(M, N, K, False, False, scale=0.3)
        bias = torch.randn(N, device="cuda")
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ref = ref_gemm(A, B, False, False) + bias
        if kw.get("aux"):
            pre = torch.empty_like(out)
            ops.gemm(A, B, bias=bias, act="gelu", aux_out=pre, out=out)
            r1 = _report(nm + ".pre", pre, ref)
            ref = gelu_tanh(ref)
            oks.append(r1["ok"])
        elif kw.get("resid"):
            res = torch.randn(M, N, device="cuda").bfloat16()
            ops.gemm(A, B, bias=bias, resid=res, out=out)
            ref = ref + res.float()
        elif kw.get("aux_in"):
            pre = torch.randn(M, N, device="cuda").bfloat16()
            ops.gemm(A, B, bias=bias, aux_in=pre, out=out)
            ref = ref * dgelu_tanh(pre.float())
        else:
            ops.gemm(A, B, bias=bias, out=out)
        r = _report(nm, out, ref)
        oks.append(r["ok"])
        rels[nm] = r["max_rel_err"]
        del A, B, out, ref
    return {"case": "staged_many_tiles", "ok": all(oks), "rels": rels}


@case("timing")
def _():
    import torch
    from multimodal_sc_b200 import ops
    res = {"case": "timing", "ok": True, "shapes": []}
    shapes = [("fc1", 262144, 1024, 256, False, False, 1), ("fc2", 262144, 256, 1024, False, False, 1),
              ("qkv", 262144, 768, 256, False, False, 1), ("proj", 262144, 256, 256, False, False, 1),
              ("patch", 262144, 256, 768, False, False, 1),
              ("dgrad_fc1", 262144, 256, 1024, False, True, 1), ("dgrad_fc2", 262144, 1024, 256, False, True, 1),
              ("wgrad_fc1", 1024, 256, 262144, True, True, 18), ("wgrad_proj", 256, 256, 262144, True, True, 74),
              ("square8k", 8192, 8192, 8192, False, False, 1)]
    for nm, M, N, K, a_mn, b_mn, ks in shapes:
        A, B = _mk(M, N, K, a_mn, b_mn)
        if ks > 1:
            out = torch.zeros(M, N, device="cuda")
            kw = dict(out_f32=out, atomic=True, ksplit=ks)
        else:
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            kw = dict(out=out)
        for _ in range(3):
            ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, **kw)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        ev0.record()
        for _ in range(n):
            ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, **kw)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / n
        flops = 2.0 * M * N * K
        byts = 2.0 * (M * K + N * K) + (4.0 if ks > 1 else 2.0) * M * N
        # cuBLAS comparison (library baseline, not the product)
        a = A.t() if a_mn else A
        b = B if b_mn else B.t()
        for _ in range(3):
            torch.matmul(a, b)
        ev0.record()
        for _ in range(n):
            torch.matmul(a, b)
        ev1.record()
        torch.cuda.synchronize()
        ms_cublas = ev0.elapsed_time(ev1) / n
        ms_generic = None
        if ks == 1:
            ops.debug_set("generic_epi", 1)
            for _ in range(2):
                ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, **kw)
            ev0.record()
            for _ in range(n):
                ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, **kw)
            ev1.record()
            torch.cuda.synchronize()
            ms_generic = ev0.elapsed_time(ev1) / n
            ops.debug_set("generic_epi", 0)
        ms_other = {}
        for label, key, val, restore in (("nwg2", "nwg", 2, int(os.environ.get("SC_NWG", "1"))),
                                         ("cta2", "cta2", 1, 1 if os.environ.get("SC_CTA2") else 0),
                                         ("cta2_nwg2", "both", 0, 0)):
            if key == "both":
                ops.debug_set("cta2", 1); ops.debug_set("nwg", 2)
            else:
                ops.debug_set(key, val)
            for _ in range(2):
                ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, **kw)
            ev0.record()
            for _ in range(n):
                ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, **kw)
            ev1.record()
            torch.cuda.synchronize()
            ms_other[label] = ev0.elapsed_time(ev1) / n
            ops.debug_set("cta2", 1 if os.environ.get("SC_CTA2") else 0)
            ops.debug_set("nwg", int(os.environ.get("SC_NWG", "1")))
        res["shapes"].append({"name": nm, "ms_generic_epi": ms_generic, "ms_other_cta_mode": ms_other, "M": M, "N": N, "K": K, "ms": ms, "tflops": flops / ms / 1e9,
                              "gbs": byts / ms / 1e6, "ms_cublas": ms_cublas})
        del A, B, out
    return res


def run_case(name):
    """returns (ok, fatal): fatal = the CUDA context is unusable, restart the worker."""
    import torch
    os.makedirs("gpurun_out", exist_ok=True)
    t0 = time.time()
    fatal = False
    try:
        from multimodal_sc_b200 import ops as _ops
        if os.environ.get("SC_CTA2"):
            _ops.debug_set("cta2", 1)
        if os.environ.get("SC_NWG"):
            _ops.debug_set("nwg", int(os.environ["SC_NWG"]))
        rec = CASES[name]()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        rec = {"case": name, "ok": False, "exception": repr(e)[:500]}
        fatal = True
    finally:
        try:
            from multimodal_sc_b200 import ops
            ops.debug_set("reset", 0)
        except Exception:  # noqa: BLE001
            pass
    rec["seconds"] = round(time.time() - t0, 2)
    with open(OUT, "a") as f:
        f.write(json.dumps(rec) + "\n")
    print(json.dumps(rec), flush=True)
    return bool(rec.get("ok")), fatal


def worker(names):
    """run cases in order; exit code = number of cases consumed (so the parent can resume)."""
    done = 0
    for nm in names:
        ok, fatal = run_case(nm)
        done += 1
        if fatal:
            break
    os._exit(done)          # skip CUDA teardown of a possibly broken context


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worker", nargs="*")
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--cases", nargs="*")
    a = ap.parse_args()
    if a.worker is not None:
        worker(a.worker)
    names = a.cases or list(CASES)
    i = 0
    while i < len(names):
        try:
            r = subprocess.run([sys.executable, __file__, "--worker", *names[i:]], timeout=900)
            consumed = r.returncode if 0 < r.returncode <= len(names) - i else 1
        except subprocess.TimeoutExpired:
            consumed = 1
            with open(OUT, "a") as f:
                f.write(json.dumps({"case": names[i], "ok": False, "exception": "worker timeout"}) + "\n")
        i += consumed
    recs = [json.loads(l) for l in open(OUT)] if os.path.exists(OUT) else []
    fails = [r["case"] for r in recs if not r.get("ok")]
    print(f"[probe] done: {len(recs)} records, failing: {fails}")
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
