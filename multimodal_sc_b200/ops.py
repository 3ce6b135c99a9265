"""Tensor-level wrappers over the C ABI (one Python function per `sc_*` entry point).

These do shape bookkeeping only; all arithmetic happens in libsemcom_sm100.so on the
current CUDA stream.  Inputs must already live on the GPU; nothing here falls back to torch.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import GemmEpilogue, SC_ACT, SC_CH, check, ptr, stream_ptr

BF16 = torch.bfloat16
F32 = torch.float32

# ---- optional per-launch timing (bench.py's roofline pass): CUDA events on the launching stream ----
_PROF = None


class profile:
    """with ops.profile() as prof: ...   -> prof.records = [(kernel, ms, flops, bytes)]"""

    def __enter__(self):
        global _PROF
        self.pending = []
        self.records = []
        _PROF = self
        return self

    def __exit__(self, *exc):
        global _PROF
        _PROF = None
        torch.cuda.synchronize()
        self.records = [(n, e0.elapsed_time(e1), fl, by) for n, e0, e1, fl, by in self.pending]
        self.pending = []
        return False

    def summary(self):
        out = {}
        for n, ms, fl, by in self.records:
            d = out.setdefault(n, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += ms
            d["flops"] += fl
            d["bytes"] += by
        return out


def _profiled(name):
    def deco(fn):
        def wrapper(*a, **k):
            if _PROF is None:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            nm, fl, by = name, 0.0, 0.0
            if name == "gemm":
                A, B = a[0], a[1]
                a_mn, b_mn = k.get("a_mn", False), k.get("b_mn", False)
                K, M = (A.shape if a_mn else A.shape[::-1])
                N = B.shape[1] if b_mn else B.shape[0]
                nm = "gemm_" + ("wgrad" if a_mn else ("dgrad" if b_mn else "fwd"))
                if getattr(_PROF, "detail", False):
                    nm += f"[{M}x{N}x{K}]"
                fl = 2.0 * M * N * K
                by = 2.0 * (M * K + N * K) + (4.0 if k.get("out_f32") is not None and k.get("out") is None else 2.0) * M * N
                for extra in ("resid", "aux_in", "aux_out", "ln"):
                    if k.get(extra) is not None:
                        by += 2.0 * M * N
            elif name in ("mlp_fwd", "mlp_bwd_dgrad"):
                M, D = a[0].shape
                R = a[1].shape[0]
                fl = 2.0 * 2.0 * M * D * R
                if name == "mlp_fwd":       # x + resid + out, hpre + h written, weights once
                    by = 2.0 * M * D * (2 + (k.get("resid") is not None)) + 2.0 * M * R * ((k.get("hpre") is not None) + (k.get("h") is not None)) + 4.0 * D * R
                else:                       # dy + dx, hpre read, dhpre written
                    by = 2.0 * M * D * 2 + 2.0 * M * R * 2 + 4.0 * D * R
            elif name in ("attn_fwd", "attn_bwd"):
                qkv, Bn, S, H = a[0], a[-3] if name == "attn_bwd" else a[1], a[-2] if name == "attn_bwd" else a[2], a[-1] if name == "attn_bwd" else a[3]
                Sk = 128            # every item is a 128-row tile (two sequences share one when S = 64)
                items = (Bn * S + 127) // 128 * H
                fl = items * (2 if name == "attn_fwd" else 5) * 2.0 * 128 * Sk * 64
                by = qkv.numel() * 2.0 * (4.0 / 3.0 if name == "attn_fwd" else 7.0 / 3.0)
            else:
                by = float(_ALGO_BYTES.get(name, lambda a, k, r: 0.0)(a, k, r))
            _PROF.pending.append((nm, e0, e1, fl, by))
            return r
        wrapper.__name__ = fn.__name__
        wrapper.__doc__ = fn.__doc__
        return wrapper
    return deco


def _nbytes(*ts):
    return sum(t.numel() * t.element_size() for t in ts if t is not None)


# algorithmic bytes of the memory-bound kernels (each operand read once, each result written once) for the roofline
# pass of bench.py: name -> f(args, kwargs, result)
_ALGO_BYTES = {
    "patchify": lambda a, k, r: (r.numel() * (2 if a[0].dtype == BF16 else 4)) + _nbytes(r),
    "layernorm_fwd": lambda a, k, r: _nbytes(a[0]) * 2 + a[0].numel() // a[0].shape[-1] * 8,
    "layernorm_bwd": lambda a, k, r: _nbytes(a[0], a[1], k.get("dres"), r) + a[0].numel() // a[0].shape[-1] * 8,
    "colsum": lambda a, k, r: _nbytes(a[0]),
    "pnorm_channel_fwd": lambda a, k, r: (k["z_strides"][2] * 2 * k["z_strides"][3] * 4 if k.get("z_strides") else _nbytes(a[0]))
    + _nbytes(k.get("y_f32"), k.get("y_bf16"), k.get("noise")),
    "pnorm_channel_bwd": lambda a, k, r: _nbytes(a[0]) * 2 + _nbytes(k.get("dz_f32"), k.get("dz_bf16")),
    "meanpool_fwd": lambda a, k, r: _nbytes(a[0]),
    "mse_fwd_bwd": lambda a, k, r: _nbytes(a[0]) * 2 + _nbytes(k.get("dr")),
    "cast_f32_bf16": lambda a, k, r: _nbytes(a[0]) * 3 // 2,
    "heads_fwd": lambda a, k, r: _nbytes(a[0], a[1], r),
    "heads_bwd": lambda a, k, r: _nbytes(a[0], a[1], a[2]) + _nbytes(a[1]),
    "sample_action": lambda a, k, r: _nbytes(a[0], a[1]) + a[1].numel() * 12,
    "gae": lambda a, k, r: a[0].numel() * 20 + _nbytes(a[3]),
    "adv_normalize": lambda a, k, r: _nbytes(a[0]) * 2,
    "ppo_loss": lambda a, k, r: _nbytes(a[0]) * 2 + a[0].shape[0] * 20,
    "adam_flat": lambda a, k, r: a[0].numel() * 30,
    "step_begin": lambda a, k, r: _nbytes(a[0]),
    "gather_rows": lambda a, k, r: 2 * _nbytes(a[2]),
}


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.SemcomError("semcom ops need CUDA tensors (no CPU fallback)")


@_profiled("gemm")
def gemm(A, B, *, a_mn=False, b_mn=False, bias=None, rowadd=None, rowgrp=None, rowgrp_size=0,
         rowgrp_scale=1.0, alpha=1.0, resid=None, aux_in=None, aux_out=None, out=None,
         out_f32=None, act=None, atomic=False, ksplit=1, colsum=None, ln=None, aux_in_erf=False):
    """acc[m,n] = sum_k A(m,k) B(n,k) with the fused epilogue of sc_gemm_bf16.

    a_mn=False: A is [M,K]; a_mn=True: A is [K,M].  b_mn=False: B is [N,K]; b_mn=True: B is [K,N].
    `out` (bf16 [M,N]) and/or `out_f32` (fp32 [M,N]) receive the result; both may be strided
    views with the same leading dimension.
    `ln` = dict(gamma, beta, eps, y[, mean, rstd]) fuses LayerNorm(out) as a second bf16 output (N == 256 only)."""
    _cuda(A, B)
    assert A.dtype == BF16 and B.dtype == BF16 and A.dim() == 2 and B.dim() == 2
    assert A.stride(1) == 1 and B.stride(1) == 1
    if a_mn:
        K, M = A.shape
    else:
        M, K = A.shape
    if b_mn:
        Kb, N = B.shape
    else:
        N, Kb = B.shape
    assert K == Kb, f"K mismatch {K} vs {Kb}"
    o = out if out is not None else out_f32
    assert o is not None and o.shape == (M, N) and o.stride(1) == 1
    ldo = o.stride(0)
    for t in (out, out_f32, resid, aux_in, aux_out):
        if t is not None:
            assert t.shape == (M, N) and t.stride(1) == 1 and t.stride(0) == ldo
    if out is not None:
        assert out.dtype == BF16
    if out_f32 is not None:
        assert out_f32.dtype == F32
    rowadd_f32 = rowadd if (rowadd is not None and rowadd.dtype == F32) else None
    rowadd_b16 = rowadd if (rowadd is not None and rowadd.dtype == BF16) else None
    epi = GemmEpilogue(
        bias=ptr(bias), rowadd=ptr(rowadd_f32), rowadd_bf16=ptr(rowadd_b16),
        rowadd_period=(rowadd.shape[0] if rowadd is not None else 0),
        rowgrp=ptr(rowgrp), rowgrp_size=rowgrp_size, rowgrp_scale=rowgrp_scale, alpha=alpha,
        resid=ptr(resid), aux_in=ptr(aux_in), aux_out=ptr(aux_out), out_bf16=ptr(out),
        out_f32=ptr(out_f32), ldo=ldo, act=SC_ACT[act], atomic=int(atomic), colsum=ptr(colsum),
        aux_in_erf=int(bool(aux_in_erf)))
    if ln is not None:
        y = ln["y"]
        assert y.dtype == BF16 and y.shape == (M, N) and y.stride(1) == 1 and y.stride(0) == ldo
        epi.ln_out, epi.ln_gamma, epi.ln_beta = ptr(y), ptr(ln["gamma"]), ptr(ln["beta"])
        epi.ln_mean, epi.ln_rstd, epi.ln_eps = ptr(ln.get("mean")), ptr(ln.get("rstd")), float(ln["eps"])
    rc = _lib.load().sc_gemm_bf16(ptr(A), ptr(B), M, N, K, int(a_mn), int(b_mn), A.stride(0),
                                  B.stride(0), C.byref(epi), ksplit, stream_ptr(A))
    check(rc, "sc_gemm_bf16")
    return o


@_profiled("mlp_fwd")
def mlp_fwd(x, w1, b1, w2, b2, *, resid=None, hpre=None, h=None, out=None, ln=None):
    """out = resid + gelu(x w1^T + b1) w2^T + b2 in one kernel (sc_mlp_fwd).  x [M,256], w1 [R,256], w2 [256,R] bf16;
    hpre / h (bf16 [M,R]) are written when given (training).  ln (dict gamma / beta / eps / y [/ mean / rstd], inference
    form only): y = LayerNorm(out) is written by the same kernel (sc_mlp_fwd_ln)."""
    _cuda(x, w1, w2)
    M, D = x.shape
    R = w1.shape[0]
    assert x.dtype == BF16 and w1.dtype == BF16 and w2.dtype == BF16 and x.is_contiguous()
    assert w1.shape == (R, D) and w2.shape == (D, R) and w1.is_contiguous() and w2.is_contiguous()
    if out is None:
        out = torch.empty(M, D, device=x.device, dtype=BF16)
    for t, shp in ((resid, (M, D)), (out, (M, D)), (hpre, (M, R)), (h, (M, R))):
        if t is not None:
            assert t.dtype == BF16 and t.shape == shp and t.is_contiguous()
    if ln is not None:
        assert hpre is None and h is None, "the fused LayerNorm exists for the inference form only"
        y = ln["y"]
        assert y.dtype == BF16 and y.shape == (M, D) and y.is_contiguous() and ln["gamma"].dtype == F32 and ln["beta"].dtype == F32
        for t in (ln.get("mean"), ln.get("rstd")):
            assert t is None or (t.dtype == F32 and t.numel() == M and t.is_contiguous())
        rc = _lib.load().sc_mlp_fwd_ln(ptr(x), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(resid), ptr(out), ptr(ln["gamma"]),
                                       ptr(ln["beta"]), float(ln["eps"]), ptr(y), ptr(ln.get("mean")), ptr(ln.get("rstd")),
                                       M, D, R, stream_ptr(x))
        check(rc, "sc_mlp_fwd_ln")
        return out
    rc = _lib.load().sc_mlp_fwd(ptr(x), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(resid), ptr(hpre), ptr(h), ptr(out),
                                M, D, R, stream_ptr(x))
    check(rc, "sc_mlp_fwd")
    return out


@_profiled("mlp_bwd_dgrad")
def mlp_bwd_dgrad(dy, w1, w2, hpre, *, dhpre, dx):
    """dhpre = (dy w2) * gelu'(hpre), dx = dhpre w1 in one kernel (sc_mlp_bwd_dgrad)."""
    _cuda(dy, w1, w2, hpre, dhpre, dx)
    M, D = dy.shape
    R = w1.shape[0]
    for t, shp in ((dy, (M, D)), (dx, (M, D)), (hpre, (M, R)), (dhpre, (M, R)), (w1, (R, D)), (w2, (D, R))):
        assert t.dtype == BF16 and t.shape == shp and t.is_contiguous()
    rc = _lib.load().sc_mlp_bwd_dgrad(ptr(dy), ptr(w1), ptr(w2), ptr(hpre), ptr(dhpre), ptr(dx), M, D, R, stream_ptr(dy))
    check(rc, "sc_mlp_bwd_dgrad")
    return dx


def debug_set(key: str, value: int):
    check(_lib.load().sc_debug_set(key.encode(), value), "sc_debug_set")


@_profiled("patchify")
def patchify(img, patch: int, out=None, row_idx=None):
    """img fp32 or bf16 [*,C,H,W] -> bf16 [B*(H/P)*(W/P), C*P*P].  row_idx (int64 [B], optional): output sample b is
    image row_idx[b] (fused minibatch gather); B = len(row_idx)."""
    _cuda(img, row_idx)
    assert img.dtype in (F32, BF16) and img.is_contiguous()
    B, Cc, H, W = img.shape
    if row_idx is not None:
        assert row_idx.dtype == torch.int64 and row_idx.is_contiguous() and row_idx.dim() == 1
        B = row_idx.numel()
    T = (H // patch) * (W // patch)
    if out is None:
        out = torch.empty(B * T, Cc * patch * patch, device=img.device, dtype=BF16)
    assert out.numel() == B * T * Cc * patch * patch
    check(_lib.load().sc_patchify_bf16(ptr(img), int(img.dtype == BF16), ptr(out), B, Cc, H, W, patch, ptr(row_idx),
                                       stream_ptr(img)), "sc_patchify_bf16")
    return out


@_profiled("layernorm_fwd")
def layernorm_fwd(x, gamma, beta, eps, y=None, mean=None, rstd=None):
    _cuda(x, gamma, beta)
    assert x.is_contiguous() and x.dtype in (BF16, F32)
    D = x.shape[-1]
    M = x.numel() // D
    if y is None:
        y = torch.empty_like(x)
    check(_lib.load().sc_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), M, D, eps,
                                       int(x.dtype == F32), stream_ptr(x)), "sc_layernorm_fwd")
    return y


@_profiled("layernorm_bwd")
def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, dres=None, dx=None, colsum=None):
    _cuda(dy, x)
    assert dy.is_contiguous() and x.is_contiguous() and dy.dtype == x.dtype
    D = x.shape[-1]
    M = x.numel() // D
    if dx is None:
        dx = torch.empty_like(x)
    check(_lib.load().sc_layernorm_bwd(ptr(dy), ptr(x), ptr(mean), ptr(rstd), ptr(gamma), ptr(dres), ptr(dx),
                                       ptr(dgamma), ptr(dbeta), ptr(colsum), M, D, int(x.dtype == F32),
                                       stream_ptr(x)), "sc_layernorm_bwd")
    return dx


@_profiled("colsum")
def colsum(x, out):
    """out[n] += sum_m x[m, n]; x bf16 [M,N] (row stride allowed)"""
    _cuda(x, out)
    assert x.dtype == BF16 and x.dim() == 2 and x.stride(1) == 1 and out.dtype == F32
    M, N = x.shape
    check(_lib.load().sc_colsum_bf16(ptr(x), ptr(out), M, N, x.stride(0), stream_ptr(x)), "sc_colsum_bf16")
    return out


@_profiled("pnorm_channel_fwd")
def pnorm_channel_fwd(z, kind, snr_db, *, seed=0, sample_offset=0, noise=None, fading=None, eps=1e-12,
                      y_f32=None, y_bf16=None, scale=None, z_strides=None, y_strides=None, seed_dev=None,
                      rician_k=0.0, sample_idx=None, quant_bits=0, quant_clip=4.0):
    """z fp32 logical [B,2,k] addressed with z_strides=(sb, sm) (default: contiguous [B,2,k]).
    `seed_dev` (device int64 [1], optional) is added to `seed` inside the kernel (CUDA-graph replays).
    Returns (y_f32, y_bf16, scale)."""
    _cuda(z)
    assert z.dtype == F32
    if z_strides is None:
        B, _, k = z.shape
        z_strides = (z.stride(0), z.stride(1))
    else:
        B, k = z_strides[2], z_strides[3]
        z_strides = z_strides[:2]
    if y_strides is None:
        y_strides = (2 * k, k)
    if y_f32 is None and y_bf16 is None:
        y_f32 = torch.empty(B, 2, k, device=z.device, dtype=F32)
    if scale is None:
        scale = torch.empty(B, 2, device=z.device, dtype=F32)
    snr_t, snr_s = (snr_db, 0.0) if torch.is_tensor(snr_db) else (None, float(snr_db))
    if snr_t is not None:
        assert snr_t.dtype == F32 and snr_t.numel() == B and snr_t.is_cuda
    check(_lib.load().sc_pnorm_channel_fwd(ptr(z), z_strides[0], z_strides[1], ptr(y_f32), ptr(y_bf16),
                                           y_strides[0], y_strides[1], ptr(scale), B, k, SC_CH[kind],
                                           float(rician_k), ptr(snr_t), snr_s, int(seed) & (2 ** 64 - 1), sample_offset,
                                           ptr(_idx(sample_idx, B)), ptr(noise), ptr(fading), eps, ptr(seed_dev),
                                           int(quant_bits), float(quant_clip), stream_ptr(z)),
          "sc_pnorm_channel_fwd")
    return y_f32, y_bf16, scale


@_profiled("pnorm_channel_bwd")
def pnorm_channel_bwd(dy, z, scale, kind, snr_db, *, seed=0, sample_offset=0, fading=None, dz_f32=None,
                      dz_bf16=None, z_strides=None, y_strides=None, seed_dev=None, rician_k=0.0, sample_idx=None,
                      quant_bits=0, quant_clip=4.0):
    _cuda(dy, z, scale)
    assert dy.dtype == F32 and z.dtype == F32
    if z_strides is None:
        B, _, k = z.shape
        z_strides = (z.stride(0), z.stride(1))
    else:
        B, k = z_strides[2], z_strides[3]
        z_strides = z_strides[:2]
    if y_strides is None:
        y_strides = (2 * k, k)
    if dz_f32 is None and dz_bf16 is None:
        dz_f32 = torch.empty(B, 2, k, device=z.device, dtype=F32)
    snr_t, snr_s = (snr_db, 0.0) if torch.is_tensor(snr_db) else (None, float(snr_db))
    check(_lib.load().sc_pnorm_channel_bwd(ptr(dy), y_strides[0], y_strides[1], ptr(z), z_strides[0],
                                           z_strides[1], ptr(scale), ptr(dz_f32), ptr(dz_bf16), B, k,
                                           SC_CH[kind], float(rician_k), ptr(snr_t), snr_s, int(seed) & (2 ** 64 - 1),
                                           sample_offset, ptr(_idx(sample_idx, B)), ptr(fading), ptr(seed_dev),
                                           int(quant_bits), float(quant_clip), stream_ptr(z)),
          "sc_pnorm_channel_bwd")
    return dz_f32, dz_bf16


def _idx(t, n):
    if t is not None:
        assert t.is_cuda and t.dtype == torch.int64 and t.is_contiguous() and t.numel() == n, "index vector: cuda int64 [n]"
    return t


def philox_normals(n_quads, seed, quad_offset=0, exact=False, device="cuda"):
    """normals of the in-kernel generator (stream 0): fp32 [n_quads, 4]"""
    out = torch.empty(n_quads, 4, device=device, dtype=F32)
    check(_lib.load().sc_philox_normals(ptr(out), n_quads, int(seed) & (2 ** 64 - 1), quad_offset, int(exact),
                                        stream_ptr(out)), "sc_philox_normals")
    return out


def philox_words(n_quads, stream_id, seed, quad_offset=0, device="cuda"):
    out = torch.empty(n_quads, 4, device=device, dtype=torch.int32)
    check(_lib.load().sc_philox_words(ptr(out), n_quads, stream_id, int(seed) & (2 ** 64 - 1), quad_offset,
                                      stream_ptr(out)), "sc_philox_words")
    return out


@_profiled("meanpool_fwd")
def meanpool_fwd(x, out_bf16=None, out_f32=None):
    _cuda(x)
    assert x.dtype == BF16 and x.is_contiguous() and x.dim() == 3
    B, S, D = x.shape
    if out_bf16 is None and out_f32 is None:
        out_bf16 = torch.empty(B, D, device=x.device, dtype=BF16)
    check(_lib.load().sc_meanpool_fwd(ptr(x), ptr(out_bf16), ptr(out_f32), B, S, D, stream_ptr(x)),
          "sc_meanpool_fwd")
    return out_bf16 if out_bf16 is not None else out_f32


@_profiled("mse_fwd_bwd")
def mse_fwd_bwd(r, t, loss_accum, dr=None, grad_scale=1.0, t1=None, seg_rows=0):
    """loss_accum[0] += mean((r-t)^2); dr = bf16(2 (r-t) grad_scale).
    With t1: r is [samples, 2*seg_rows, D]; t / t1 are the per-modality targets [samples*seg_rows, D]."""
    _cuda(r, t, loss_accum)
    assert r.dtype == BF16 and t.dtype == BF16 and r.is_contiguous() and t.is_contiguous()
    n = r.numel()
    D = r.shape[-1]
    check(_lib.load().sc_mse_fwd_bwd(ptr(r), ptr(t), ptr(t1), seg_rows, D, ptr(loss_accum), ptr(dr), n, 1.0 / n,
                                     grad_scale, stream_ptr(r)), "sc_mse_fwd_bwd")
    return dr


@_profiled("cast_f32_bf16")
def cast_f32_bf16(src, dst=None):
    _cuda(src)
    assert src.dtype == F32 and src.is_contiguous()
    if dst is None:
        dst = torch.empty_like(src, dtype=BF16)
    check(_lib.load().sc_cast_f32_bf16(ptr(src), ptr(dst), src.numel(), stream_ptr(src)), "sc_cast_f32_bf16")
    return dst


@_profiled("attn_fwd")
def attn_fwd(qkv, B, S, H, out=None, lse=None):
    """qkv bf16 [B*S, 3*H*64] -> out bf16 [B*S, H*64], lse fp32 [B*S, H] (log2 domain)"""
    _cuda(qkv)
    assert qkv.dtype == BF16 and qkv.is_contiguous() and qkv.shape == (B * S, 3 * H * 64)
    if out is None:
        out = torch.empty(B * S, H * 64, device=qkv.device, dtype=BF16)
    if lse is None:
        lse = torch.empty(B * S, H, device=qkv.device, dtype=F32)
    check(_lib.load().sc_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, S, H, stream_ptr(qkv)), "sc_attn_fwd")
    return out, lse


@_profiled("attn_bwd")
def attn_bwd(qkv, out, dout, lse, B, S, H, dqkv=None):
    _cuda(qkv, dout, lse)
    assert qkv.dtype == BF16 and dout.dtype == BF16 and qkv.is_contiguous() and dout.is_contiguous()
    if dqkv is None:
        dqkv = torch.empty_like(qkv)
    check(_lib.load().sc_attn_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(dqkv), B, S, H, stream_ptr(qkv)),
          "sc_attn_bwd")
    return dqkv


@_profiled("heads_fwd")
def heads_fwd(trunk, W, bias, out=None):
    """trunk bf16 [B,D], W fp32 [A1,D], bias fp32 [A1] -> fp32 [B,A1]"""
    _cuda(trunk, W, bias)
    B, D = trunk.shape
    A1 = W.shape[0]
    assert trunk.dtype == BF16 and W.dtype == F32 and W.is_contiguous() and trunk.is_contiguous()
    if out is None:
        out = torch.empty(B, A1, device=trunk.device, dtype=F32)
    check(_lib.load().sc_heads_fwd(ptr(trunk), ptr(W), ptr(bias), ptr(out), B, D, A1, stream_ptr(trunk)), "sc_heads_fwd")
    return out


@_profiled("heads_bwd")
def heads_bwd(dho, trunk, W, dW, dbias, dtrunk_pre=None):
    _cuda(dho, trunk, W, dW, dbias)
    B, D = trunk.shape
    A1 = W.shape[0]
    assert dho.dtype == F32 and dho.is_contiguous() and dho.shape == (B, A1)
    if dtrunk_pre is None:
        dtrunk_pre = torch.empty_like(trunk)
    check(_lib.load().sc_heads_bwd(ptr(dho), ptr(trunk), ptr(W), ptr(dtrunk_pre), ptr(dW), ptr(dbias), B, D, A1,
                                   stream_ptr(dho)), "sc_heads_bwd")
    return dtrunk_pre


@_profiled("sample_action")
def sample_action(head_out, u, n_actions, action=None, logp=None):
    _cuda(head_out, u)
    B, ld = head_out.shape
    assert head_out.dtype == F32 and head_out.is_contiguous() and u.dtype == F32
    if action is None:
        action = torch.empty(B, device=u.device, dtype=torch.int64)
    if logp is None:
        logp = torch.empty(B, device=u.device, dtype=F32)
    check(_lib.load().sc_sample_action(ptr(head_out), ld, ptr(u), ptr(action), ptr(logp), B, n_actions,
                                       stream_ptr(head_out)), "sc_sample_action")
    return action, logp


@_profiled("gae")
def gae(rewards, values, dones, last_value, gamma, lam, adv=None, ret=None, sums=None):
    _cuda(rewards, values, dones, last_value)
    T, N = rewards.shape
    for t in (rewards, values, dones, last_value):
        assert t.dtype == F32 and t.is_contiguous()
    if adv is None:
        adv = torch.empty_like(rewards)
    if ret is None:
        ret = torch.empty_like(rewards)
    if sums is not None:
        assert sums.dtype == torch.float64 and sums.numel() == 2
    check(_lib.load().sc_gae(ptr(rewards), ptr(values), ptr(dones), ptr(last_value), ptr(adv), ptr(ret), ptr(sums),
                             T, N, gamma, lam, stream_ptr(rewards)), "sc_gae")
    return adv, ret


@_profiled("adv_normalize")
def adv_normalize(adv, sums, count, eps):
    _cuda(adv, sums)
    assert adv.dtype == F32 and adv.is_contiguous() and sums.dtype == torch.float64
    check(_lib.load().sc_adv_normalize(ptr(adv), ptr(sums), count, adv.numel(), eps, stream_ptr(adv)),
          "sc_adv_normalize")
    return adv


@_profiled("ppo_loss")
def ppo_loss(head_out, actions, logp_old, adv, ret, n_actions, clip_eps, c_v, c_ent, inv_m, dho=None, stats=None,
             idx=None):
    """fused PPO loss fwd+bwd. stats[0:5] += (pi, v, ent, kl, clipfrac) * inv_m.  idx (int64 [M], optional): sample i
    reads actions / logp_old / adv / ret at idx[i] (fused minibatch gather).  An action outside [0, n_actions) turns
    the statistics into NaN (the kernel does not trap)."""
    _cuda(head_out, actions, logp_old, adv, ret, idx)
    M, ld = head_out.shape
    assert head_out.dtype == F32 and head_out.is_contiguous() and actions.dtype == torch.int64 and actions.is_contiguous()
    for t in (logp_old, adv, ret):
        assert t.dtype == F32 and t.is_contiguous(), "logp_old / adv / ret must be contiguous fp32"
    n_src = actions.numel()
    assert logp_old.numel() == n_src and adv.numel() == n_src and ret.numel() == n_src
    assert idx is not None or n_src == M, "without idx the per-sample vectors must have M entries"
    if dho is None:
        dho = torch.empty_like(head_out)
    if stats is None:
        stats = torch.zeros(8, device=head_out.device, dtype=F32)
    check(_lib.load().sc_ppo_loss(ptr(head_out), ld, ptr(actions), ptr(logp_old), ptr(adv), ptr(ret), ptr(_idx(idx, M)),
                                  ptr(dho), ptr(stats), M, n_actions, clip_eps, c_v, c_ent, inv_m, stream_ptr(head_out)),
          "sc_ppo_loss")
    return dho, stats


@_profiled("adam_flat")
def adam_flat(p, g, m, v, p_bf16, lr, beta1, beta2, eps, step, grad_scale=1.0, step_dev=None):
    """`step_dev` (device int64 [1], optional): effective step = step + step_dev[0], read inside the kernel."""
    _cuda(p, g, m, v)
    n = p.numel()
    check(_lib.load().sc_adam_flat(ptr(p), ptr(g), ptr(m), ptr(v), ptr(p_bf16), n, lr, beta1, beta2, eps, step,
                                   grad_scale, ptr(step_dev), stream_ptr(p)), "sc_adam_flat")


@_profiled("step_begin")
def step_begin(grad, stats, step_dev=None):
    """zero the flat gradient bucket and the statistics vector, advance the device-side step counter (one launch)"""
    _cuda(grad, stats)
    assert grad.dtype == F32 and grad.is_contiguous() and stats.dtype == F32 and stats.is_contiguous()
    if step_dev is not None:
        assert step_dev.dtype == torch.int64 and step_dev.is_cuda
    check(_lib.load().sc_step_begin(ptr(grad), grad.numel(), ptr(stats), stats.numel(), ptr(step_dev), stream_ptr(grad)),
          "sc_step_begin")


@_profiled("gather_rows")
def gather_rows(src, idx, out, max_ctas: int = 0):
    """out[i] = src[idx[i]] over the leading dimension (idx int64 on the device, or None = copy).  `src` may live in PINNED
    host memory: the kernel then pulls exactly the minibatch's rows over PCIe (PPOAgent.update's host -> device path)."""
    _cuda(out)
    if not (src.is_cuda or src.is_pinned()):
        raise _lib.SemcomError("gather_rows: the source must be device or pinned host memory")
    assert src.is_contiguous() and out.is_contiguous() and src.dtype == out.dtype and src.shape[1:] == out.shape[1:]
    n = out.shape[0]
    if idx is not None:
        assert idx.is_cuda and idx.dtype == torch.int64 and idx.numel() == n
    row_bytes = src[0].numel() * src.element_size()
    check(_lib.load().sc_gather_rows(src.data_ptr(), src.shape[0], row_bytes, ptr(idx), out.data_ptr(), n, max_ctas,
                                     stream_ptr(out)), "sc_gather_rows")
    return out


@_profiled("snr_embed_fwd")
def snr_embed_fwd(snr_db, u, v, W, B, emb=None, h=None):
    """emb[b] = W tanh(s_b u + v), s_b = snr_db[b] / 20 (tensor [B] or float).  Returns (emb fp32 [B,D], h fp32 [B,E])."""
    _cuda(u, v, W)
    D, E = W.shape
    snr_t, snr_s = (snr_db, 0.0) if torch.is_tensor(snr_db) else (None, float(snr_db))
    if snr_t is not None:
        assert snr_t.is_cuda and snr_t.dtype == F32 and snr_t.numel() == B
    if emb is None:
        emb = torch.empty(B, D, device=W.device, dtype=F32)
    if h is None:
        h = torch.empty(B, E, device=W.device, dtype=F32)
    check(_lib.load().sc_snr_embed_fwd(ptr(snr_t), snr_s, ptr(u), ptr(v), ptr(W), ptr(emb), ptr(h), B, D, E,
                                       stream_ptr(W)), "sc_snr_embed_fwd")
    return emb, h


@_profiled("snr_embed_bwd")
def snr_embed_bwd(demb, h, snr_db, W, dW, du, dv, demb_scale=1.0):
    _cuda(demb, h, W, dW, du, dv)
    B, D = demb.shape
    E = h.shape[1]
    assert demb.dtype == F32 and demb.is_contiguous() and h.is_contiguous() and W.shape == (D, E)
    snr_t, snr_s = (snr_db, 0.0) if torch.is_tensor(snr_db) else (None, float(snr_db))
    check(_lib.load().sc_snr_embed_bwd(ptr(demb), ptr(h), ptr(snr_t), snr_s, ptr(W), ptr(dW), ptr(du), ptr(dv),
                                       float(demb_scale), B, D, E, stream_ptr(W)), "sc_snr_embed_bwd")
