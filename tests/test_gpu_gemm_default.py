"""Default-mode sc_gemm_bf16 against an fp32 matmul of the same bf16 operands: all three operand-major pairs of the
model (fwd K/K, dgrad K/MN, wgrad MN/MN) x the production K and N x every epilogue kind x split-K.

Tolerance: |got - ref| <= 1e-2 * max|ref| for bf16 outputs (one bf16 rounding of the result = 2^-9 of its magnitude,
the rest is accumulation order); 1e-3 for fp32 outputs (operands are exact bf16, accumulation is fp32)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

KS = [16, 64, 256, 768, 1024]
NS = [16, 256, 768, 1024]
M_ROWS = 128 * 5 + 40                      # several row tiles and a ragged tail


def _mk(M, N, K, a_mn, b_mn, seed=0, scale=0.3):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = (torch.randn((K, M) if a_mn else (M, K), generator=g, device="cuda") * scale).bfloat16()
    B = (torch.randn((K, N) if b_mn else (N, K), generator=g, device="cuda") * scale).bfloat16()
    return A, B


def _ref(A, B, a_mn, b_mn):
    a = A.float().t() if a_mn else A.float()
    b = B.float().t() if b_mn else B.float()
    return a @ b.t()


def _check(got, ref, tol, what):
    err = float((got.float() - ref).abs().max())
    scale = float(ref.abs().max()) + 1e-30
    assert err <= tol * scale, (what, err / scale)
    assert not torch.isnan(got.float()).any(), what


def _gelu(x):
    return torch.nn.functional.gelu(x, approximate="tanh")


def _dgelu(x, approximate="tanh"):
    x = x.detach().clone().requires_grad_(True)
    (g,) = torch.autograd.grad(torch.nn.functional.gelu(x, approximate=approximate).sum(), x)
    return g


@pytest.mark.parametrize("majors", [(False, False), (False, True), (True, True)], ids=["fwd", "dgrad", "wgrad"])
@pytest.mark.parametrize("K", KS)
@pytest.mark.parametrize("N", NS)
def test_plain_all_majors(cuda_lib, majors, K, N):
    from multimodal_sc_b200 import ops
    a_mn, b_mn = majors
    M = M_ROWS if not a_mn else 256          # wgrad: M is a weight dimension
    Kr = K if not a_mn else K * 8 + 24       # wgrad: the reduction runs over tokens (long, ragged)
    A, B = _mk(M, N, Kr, a_mn, b_mn, seed=K + N)
    ref = _ref(A, B, a_mn, b_mn)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out=out)
    _check(out, ref, 1e-2, "bf16 out")
    o32 = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out_f32=o32)
    _check(o32, ref, 1e-3, "fp32 out")


@pytest.mark.parametrize("K", KS)
@pytest.mark.parametrize("N", [256, 768, 1024])
@pytest.mark.parametrize("kind", ["bias", "fc1", "resid", "dgelu", "pos16", "resid_ln", "tanh", "rowgrp", "alpha_rowadd"])
def test_forward_epilogues(cuda_lib, kind, K, N):
    """every epilogue kind of the staged / fused-LayerNorm / generic paths on the fwd operand majors (dgelu: dgrad majors)"""
    from multimodal_sc_b200 import ops
    if kind == "resid_ln" and N != 256:
        pytest.skip("fused LayerNorm needs N == 256")
    M = M_ROWS if kind != "pos16" else 128 * 6       # pos table: whole periods of 64 rows
    b_mn = kind == "dgelu"
    A, B = _mk(M, N, K, False, b_mn, seed=3 * K + N)
    acc = _ref(A, B, False, b_mn)
    g = torch.Generator(device="cuda").manual_seed(7)
    bias = torch.randn(N, generator=g, device="cuda") * 0.2
    res = (torch.randn(M, N, generator=g, device="cuda") * 0.5).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    if kind == "bias":
        ops.gemm(A, B, bias=bias, out=out)
        _check(out, acc + bias, 1e-2, kind)
    elif kind == "fc1":
        pre = torch.empty_like(out)
        ops.gemm(A, B, bias=bias, act="gelu", aux_out=pre, out=out)
        _check(pre, acc + bias, 1e-2, "pre-activation")
        _check(out, _gelu(acc + bias), 1e-2, "gelu")
    elif kind == "resid":
        ops.gemm(A, B, bias=bias, resid=res, out=out)
        _check(out, acc + bias + res.float(), 1e-2, kind)
    elif kind == "dgelu":
        ops.gemm(A, B, b_mn=True, aux_in=res, out=out)
        _check(out, acc * _dgelu(res.float()), 1e-2, kind)
    elif kind == "pos16":
        pos = (torch.randn(64, N, generator=g, device="cuda") * 0.3).bfloat16()
        ops.gemm(A, B, bias=bias, rowadd=pos, out=out)
        _check(out, acc + bias + pos.float().repeat(M // 64, 1), 1e-2, kind)
    elif kind == "resid_ln":
        gamma = 1.0 + 0.1 * torch.randn(N, generator=g, device="cuda")
        beta = 0.1 * torch.randn(N, generator=g, device="cuda")
        y = torch.empty_like(out)
        mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
        ops.gemm(A, B, bias=bias, resid=res, out=out, ln=dict(gamma=gamma, beta=beta, eps=1e-5, y=y, mean=mean, rstd=rstd))
        x = acc + bias + res.float()
        _check(out, x, 1e-2, "x")
        xr = out.float()                                # the LayerNorm sees the bf16-rounded x
        ref_y = torch.nn.functional.layer_norm(xr, (N,), gamma, beta, 1e-5)
        _check(y, ref_y, 1e-2, "ln(x)")
        assert torch.allclose(mean, xr.mean(-1), rtol=1e-4, atol=1e-5)
        assert torch.allclose(rstd, torch.rsqrt(xr.var(-1, unbiased=False) + 1e-5), rtol=1e-4, atol=1e-5)
    elif kind == "tanh":
        ops.gemm(A, B, bias=bias, act="tanh", out=out)
        _check(out, torch.tanh(acc + bias), 1e-2, kind)
    elif kind == "rowgrp":
        grp = torch.randn((M + 63) // 64, N, generator=g, device="cuda")
        ops.gemm(A, B, rowgrp=grp, rowgrp_size=64, rowgrp_scale=0.25, out=out)
        _check(out, acc + 0.25 * grp.repeat_interleave(64, 0)[:M], 1e-2, kind)
    elif kind == "alpha_rowadd":
        ra = torch.randn(40, N, generator=g, device="cuda")
        o32 = torch.empty(M, N, device="cuda")
        ops.gemm(A, B, alpha=0.5, rowadd=ra, out_f32=o32)
        idx = torch.arange(M, device="cuda") % 40
        _check(o32, 0.5 * acc + ra[idx], 1e-3, kind)


@pytest.mark.parametrize("shape", [(1024, 256), (256, 1024), (768, 256), (256, 256), (256, 768), (256, 64), (16, 256), (256, 16)])
@pytest.mark.parametrize("ksplit", [1, 2, 9, 37])
def test_wgrad_split_k_with_bias_colsum(cuda_lib, shape, ksplit):
    """the model's weight-gradient shapes [N_out, K_in] = dy^T x over a long ragged token dimension, split-K with
    red.add accumulation into a non-zero bucket, bias gradient from the all-ones MMA"""
    from multimodal_sc_b200 import ops
    Nout, Kin = shape
    tokens = 128 * 37 + 8
    g = torch.Generator(device="cuda").manual_seed(Nout + Kin + ksplit)
    dy = (torch.randn(tokens, Nout, generator=g, device="cuda") * 0.3).bfloat16()
    x = (torch.randn(tokens, Kin, generator=g, device="cuda") * 0.3).bfloat16()
    base = torch.randn(Nout, Kin, generator=g, device="cuda")
    dW, db = base.clone(), torch.zeros(Nout, device="cuda")
    ops.gemm(dy, x, a_mn=True, b_mn=True, out_f32=dW, atomic=True, ksplit=ksplit, colsum=db)
    ref = dy.float().t() @ x.float()
    _check(dW - base, ref, 1e-3, "dW")
    _check(db, dy.float().sum(0), 1e-3, "db")


@pytest.mark.parametrize("kind", ["fc1", "dgelu"])
def test_erf_gelu_epilogues(cuda_lib, kind):
    """SC_ACT_GELU_ERF / aux_in_erf (SpecV0.gelu == "erf"): the exact GELU and its derivative in the direct-store epilogue"""
    from multimodal_sc_b200 import ops
    M, N, K = M_ROWS, 1024, 256
    A, B = _mk(M, N, K, False, kind == "dgelu", seed=5)
    acc = _ref(A, B, False, kind == "dgelu")
    g = torch.Generator(device="cuda").manual_seed(1)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    if kind == "fc1":
        bias = torch.randn(N, generator=g, device="cuda") * 0.2
        pre = torch.empty_like(out)
        ops.gemm(A, B, bias=bias, act="gelu_erf", aux_out=pre, out=out)
        _check(pre, acc + bias, 1e-2, "pre")
        _check(out, torch.nn.functional.gelu(acc + bias), 1e-2, "gelu erf")
    else:
        hp = (torch.randn(M, N, generator=g, device="cuda") * 1.5).bfloat16()
        ops.gemm(A, B, b_mn=True, aux_in=hp, aux_in_erf=True, out=out)
        _check(out, acc * _dgelu(hp.float(), "none"), 1e-2, "gelu' erf")


def test_tensor_map_cache_is_transparent(cuda_lib):
    """cached CUtensorMaps (keyed by every encode argument) give bit-identical results; a different pointer / shape /
    stride under the same hash slot re-encodes"""
    from multimodal_sc_b200 import ops
    outs = {}
    for on in (1, 0, 1):
        ops.debug_set("tmap_cache", on)
        res = []
        for M in (128 * 3, 128 * 3 + 8):
            A, B = _mk(M, 256, 256, False, False, seed=M)
            o = torch.empty(M, 256, device="cuda", dtype=torch.bfloat16)
            for _ in range(2):                        # second call hits the cache
                ops.gemm(A, B, out=o)
            res.append(o.clone())
        outs.setdefault("ref", res)
        for a, b in zip(outs["ref"], res):
            assert torch.equal(a, b)
    ops.debug_set("tmap_cache", 1)
