"""GPU parity: tcgen05 attention fwd/bwd, heads, sampling, GAE, PPO loss, Adam vs the oracle."""
import math

import pytest
import torch

from multimodal_sc_b200.spec import SpecV0
from oracle import ppo as oppo
from oracle.model import sample_action as o_sample

pytestmark = pytest.mark.gpu
SPEC = SpecV0()


def _scaled(got, ref, rtol, what=""):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    err = float((got - ref).abs().max())
    scale = float(ref.abs().max()) + 1e-12
    assert err <= rtol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.3e})"


def _close(got, ref, rtol, atol, what=""):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    err = (got - ref).abs()
    bad = err > atol + rtol * ref.abs()
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} bad, max err {float(err.max()):.3e}"


def _ref_attn(qkv, B, S, H):
    q, k, v = (qkv.view(B, S, 3, H, 64)[:, :, i].transpose(1, 2) for i in range(3))
    att = torch.softmax((q @ k.transpose(-1, -2)) / 8.0, dim=-1)
    return (att @ v).transpose(1, 2).reshape(B * S, H * 64)


@pytest.mark.parametrize("B,S,H", [(1, 128, 1), (3, 128, 4), (2, 64, 4), (5, 64, 2), (300, 128, 4), (301, 64, 4)])
def test_attention_fwd_bwd(cuda_lib, B, S, H):
    from multimodal_sc_b200 import ops
    g = torch.Generator().manual_seed(B * S + H)
    qkv = (torch.randn(B * S, 3 * H * 64, generator=g) * 1.5).bfloat16()
    dout = torch.randn(B * S, H * 64, generator=g).bfloat16()
    qr = qkv.float().requires_grad_(True)
    o_ref = _ref_attn(qr, B, S, H)
    (o_ref * dout.float()).sum().backward()
    out, lse = ops.attn_fwd(qkv.cuda(), B, S, H)
    dqkv = ops.attn_bwd(qkv.cuda(), out, dout.cuda(), lse, B, S, H)
    _scaled(out, o_ref, 2e-2, "attn out")
    _scaled(dqkv, qr.grad, 2e-2, "attn dqkv")
    # lse (log2 domain) against the reference
    q, k, _ = (qkv.float().view(B, S, 3, H, 64)[:, :, i].transpose(1, 2) for i in range(3))
    lse_ref = torch.logsumexp((q @ k.transpose(-1, -2)) / 8.0, dim=-1) / math.log(2.0)     # [B,H,S]
    _close(lse.view(B, S, H).permute(0, 2, 1), lse_ref, 1e-3, 1e-2, "lse")


def test_heads_fwd_bwd(cuda_lib):
    from multimodal_sc_b200 import ops
    g = torch.Generator().manual_seed(0)
    B, D, A1 = 777, 256, 10
    trunk = torch.tanh(torch.randn(B, D, generator=g)).bfloat16()
    W = torch.randn(A1, D, generator=g) * 0.1
    b = torch.randn(A1, generator=g)
    dho = torch.randn(B, A1, generator=g)
    ho = ops.heads_fwd(trunk.cuda(), W.cuda(), b.cuda())
    _close(ho, trunk.float() @ W.t() + b, 1e-5, 1e-5, "heads fwd")
    dW = torch.zeros(A1, D, device="cuda")
    db = torch.zeros(A1, device="cuda")
    dpre = ops.heads_bwd(dho.cuda(), trunk.cuda(), W.cuda(), dW, db)
    tf = trunk.float()
    _scaled(dpre, (dho @ W) * (1 - tf * tf), 1e-2, "heads dtrunk")
    _close(dW, dho.t() @ tf, 1e-4, 1e-3, "heads dW")
    _close(db, dho.sum(0), 1e-4, 1e-3, "heads db")


def test_sample_action_bit_exact(cuda_lib):
    from multimodal_sc_b200 import ops
    g = torch.Generator().manual_seed(1)
    B, A = 20000, SPEC.n_actions
    ho = torch.randn(B, A + 1, generator=g) * 3
    u = torch.rand(B, generator=g)
    u[:4] = torch.tensor([0.0, 0.99999994, 0.5, 1e-8])
    a_ref, lp_ref = o_sample(ho[:, :A], u)
    a, lp = ops.sample_action(ho.cuda(), u.cuda(), A)
    assert torch.equal(a.cpu(), a_ref)
    _close(lp, lp_ref, 1e-5, 1e-6, "logp")


@pytest.mark.parametrize("T,N", [(128, 32), (128, 256), (1, 1), (37, 5), (200, 33), (21, 16384 + 3)])
def test_gae(cuda_lib, T, N):
    from multimodal_sc_b200 import ops
    g = torch.Generator().manual_seed(T * N)
    r = torch.randn(T, N, generator=g)
    v = torch.randn(T, N, generator=g)
    d = (torch.rand(T, N, generator=g) < 1 / 16).float()
    lv = torch.randn(N, generator=g)
    adv_ref, ret_ref = oppo.gae(r, v, d, lv, SPEC.gamma, SPEC.lam)
    sums = torch.zeros(2, device="cuda", dtype=torch.float64)
    adv, ret = ops.gae(r.cuda(), v.cuda(), d.cuda(), lv.cuda(), SPEC.gamma, SPEC.lam, sums=sums)
    _close(adv, adv_ref, 1e-5, 2e-5, "adv")
    _close(ret, ret_ref, 1e-5, 2e-5, "ret")
    if T * N > 1:
        ops.adv_normalize(adv, sums, T * N, SPEC.adv_eps)
        _close(adv, oppo.normalize_adv(adv_ref, SPEC.adv_eps), 1e-4, 1e-4, "adv norm")


def test_gae_closed_form(cuda_lib):
    """constant reward, no dones, V = 0: A_t = sum_i (gamma lam)^i r"""
    from multimodal_sc_b200 import ops
    T, N = 128, 3
    r = torch.ones(T, N, device="cuda")
    z = torch.zeros(T, N, device="cuda")
    adv, _ = ops.gae(r, z, z, torch.zeros(N, device="cuda"), SPEC.gamma, SPEC.lam)
    gl = SPEC.gamma * SPEC.lam
    ref = torch.tensor([(1 - gl ** (T - t)) / (1 - gl) for t in range(T)])
    _close(adv[:, 0], ref, 1e-5, 1e-5, "closed form")


@pytest.mark.parametrize("M", [4, 4096])
def test_ppo_loss_fwd_bwd(cuda_lib, M):
    from multimodal_sc_b200 import ops
    g = torch.Generator().manual_seed(M)
    A = SPEC.n_actions
    ho = torch.randn(M, A + 1, generator=g)
    act = torch.randint(0, A, (M,), generator=g)
    lp_old = torch.log_softmax(ho[:, :A], -1).gather(-1, act[:, None]).squeeze(-1) + 0.3 * torch.randn(M, generator=g)
    adv = torch.randn(M, generator=g)
    ret = torch.randn(M, generator=g)
    hr = ho.clone().requires_grad_(True)
    out = oppo.ppo_loss(hr[:, :A], act, lp_old, adv, hr[:, A], ret, torch.tensor(0.0), SPEC)
    out["total"].backward()
    dho, stats = ops.ppo_loss(ho.cuda(), act.cuda(), lp_old.cuda(), adv.cuda(), ret.cuda(), A, SPEC.clip_eps,
                              SPEC.c_v, SPEC.c_ent, 1.0 / M)
    # fp32 rtol 1e-5 (north_star).  The logit gradient inv_m (gp (1[j=a] - p_j) + c_ent p_j (lp_j + H)) cancels O(1)
    # terms, so the element-wise floor is ulp(1) * inv_m, not ulp(result): the absolute term is 1e-5 of the tensor's scale
    scale = float(hr.grad.abs().max())
    _close(dho, hr.grad, 1e-5, 1e-5 * scale, "dho")
    s = stats.cpu()
    for i, k in enumerate(["pi", "v", "ent", "kl", "clipfrac"]):
        assert abs(float(s[i]) - float(out[k])) <= 1e-6 + 1e-5 * abs(float(out[k])), (k, float(s[i]), float(out[k]))


def test_adam_matches_torch(cuda_lib):
    from multimodal_sc_b200 import ops
    g = torch.Generator().manual_seed(5)
    n = 100003
    p0 = torch.randn(n, generator=g)
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=SPEC.lr, betas=(SPEC.beta1, SPEC.beta2), eps=SPEC.adam_eps)
    p = p0.clone().cuda()
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    pb = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    for step in range(1, 4):
        gr = torch.randn(n, generator=g)
        pr.grad = gr.clone()
        opt.step()
        ops.adam_flat(p, (gr * 2).cuda(), m, v, pb, SPEC.lr, SPEC.beta1, SPEC.beta2, SPEC.adam_eps, step,
                      grad_scale=0.5)
    _close(p, pr.detach(), 1e-5, 1e-6, "adam p")
    assert torch.equal(pb.cpu(), p.cpu().bfloat16())


@pytest.mark.parametrize("mode", [{"cta2": 1}, {"nwg": 2}, {"cta2": 1, "nwg": 2}])
def test_gemm_alternative_schedules(cuda_lib, mode):
    """the cta_group::2 (CTA-pair) and two-epilogue-warpgroup variants of the GEMM stay correct (they are kept
    as options; the default is one CTA per tile, one epilogue warpgroup, four pipeline stages)."""
    from multimodal_sc_b200 import ops
    try:
        for k, v in mode.items():
            ops.debug_set(k, v)
        g = torch.Generator(device="cuda").manual_seed(0)
        M, N, K = 128 * 301, 1024, 256
        A = (torch.randn(M, K, generator=g, device="cuda") * 0.3).bfloat16()
        W = (torch.randn(N, K, generator=g, device="cuda") * 0.3).bfloat16()
        bias = torch.randn(N, device="cuda")
        res = torch.randn(M, N, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ops.gemm(A, W, bias=bias, resid=res, out=out)
        _scaled(out, A.float() @ W.float().t() + bias + res.float(), 1e-2, "fwd+resid")
        dW = torch.zeros(N, K, device="cuda")
        db = torch.zeros(N, device="cuda")
        ops.gemm(out, A, a_mn=True, b_mn=True, out_f32=dW, atomic=True, ksplit=9, colsum=db)
        _close(dW, out.float().t() @ A.float(), 1e-3, 1e-1, "wgrad")
        _close(db, out.float().sum(0), 1e-3, 1e-1, "bias grad")
    finally:
        ops.debug_set("reset", 0)
