"""GPU parity of the whole path: native SemComPolicy / PPOAgent (bf16 tensor-core path) vs the fp32
oracle on identical weights, inputs and noise seeds.  Tolerance: bf16 rtol 2e-2 of the output
scale for activations; gradients per tensor by max-abs error relative to the tensor's scale.  These are the SMALL
cases (B <= 8, plus the schedule-switch self-consistency at B = 256); the same comparisons at BASELINE.json's real sizes
(B = 256 / 1024 / 4096, oracle in fp32 on the GPU) are in tests/test_gpu_parity_full.py."""
import pytest
import torch

from multimodal_sc_b200.spec import SpecV0
from oracle.model import SemComPolicy as OraclePolicy
from oracle.ppo import PPOAgent as OracleAgent

pytestmark = pytest.mark.gpu
SPEC = SpecV0()


def _pair(std=0.05, seed=3):
    from multimodal_sc_b200.model import SemComPolicy
    o = OraclePolicy(SPEC)
    o.reset_parameters(std=std, seed=seed)
    # non-trivial LayerNorm affine + biases so every gradient path is exercised
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in o.named_parameters():
            if n.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            elif ".ln" in n or "norm" in n:
                p.copy_(1.0 + torch.randn(p.shape, generator=g) * 0.1)
    n = SemComPolicy(SPEC)
    n.load_state_dict(o.state_dict())
    return o, n


def _inputs(B, seed=0):
    g = torch.Generator().manual_seed(seed)
    cam = torch.rand(B, 3, 128, 128, generator=g)
    bev = (torch.rand(B, 1, 64, 64, generator=g) < 0.1).float() * torch.rand(B, 1, 64, 64, generator=g)
    return cam, bev


def _rel(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return float((got - ref).abs().max() / (ref.abs().max() + 1e-12))


def test_state_dict_roundtrip(cuda_lib):
    o, n = _pair()
    sd_o, sd_n = o.state_dict(), n.state_dict()
    assert list(sd_o.keys()) == list(sd_n.keys())
    for k in sd_o:
        assert torch.equal(sd_o[k], sd_n[k].cpu()), k
    o2 = OraclePolicy(SPEC)
    o2.load_state_dict({k: v.cpu() for k, v in sd_n.items()})     # loads back the other way


@pytest.mark.parametrize("channel,B", [("awgn", 4), ("rayleigh", 3), ("rician", 3), ("none", 2)])
def test_forward_parity(cuda_lib, channel, B):
    o, n = _pair()
    cam, bev = _inputs(B)
    snr = torch.linspace(-5, 20, B)
    with torch.no_grad():
        lo, vo, ao = o(cam, bev, snr, channel=channel, seed=1234, sample_offset=11)
        ln, vn, an = n(cam.cuda(), bev.cuda(), snr.cuda(), channel=channel, seed=1234, sample_offset=11)
    assert _rel(an["z"], ao["z"]) < 2e-2, ("z", _rel(an["z"], ao["z"]))
    assert _rel(an["y"], ao["y"]) < 2e-2, ("y", _rel(an["y"], ao["y"]))
    assert _rel(ln, lo) < 2e-2, ("logits", _rel(ln, lo))
    assert _rel(vn, vo) < 2e-2, ("value", _rel(vn, vo))
    assert abs(float(an["l_sem"]) - float(ao["l_sem"])) < 2e-2 * float(ao["l_sem"]), (float(an["l_sem"]), float(ao["l_sem"]))


def test_act_matches_on_identical_logits(cuda_lib):
    """a7: integer actions are bit-exact given identical logits and uniforms."""
    from multimodal_sc_b200 import ops
    from oracle.model import sample_action
    o, n = _pair()
    cam, bev = _inputs(6)
    u = torch.rand(6)
    a, logp, v = n.act(cam.cuda(), bev.cuda(), 10.0, u.cuda(), seed=5)
    with torch.no_grad():
        ln, _, _ = n(cam.cuda(), bev.cuda(), 10.0, seed=5)
    a_ref, lp_ref = sample_action(ln.cpu(), u)
    assert torch.equal(a.cpu(), a_ref)
    assert torch.allclose(logp.cpu(), lp_ref, rtol=1e-5, atol=1e-6)


def test_backward_parity(cuda_lib):
    o, n = _pair()
    B = 4
    cam, bev = _inputs(B, seed=1)
    g = torch.Generator().manual_seed(9)
    gl, gv = torch.randn(B, SPEC.n_actions, generator=g), torch.randn(B, generator=g)

    def loss_of(model, dev):
        l, v, aux = model(cam.to(dev), bev.to(dev), 7.0, channel="awgn", seed=77)
        return (l * gl.to(dev)).sum() + (v * gv.to(dev)).sum() + 0.7 * aux["l_sem"]

    loss_of(o, "cpu").backward()
    loss_of(n, "cuda").backward()
    # every gradient tensor: max-abs error relative to the tensor's scale.  At B = 4 each gradient is a sum over very few
    # samples, so the bf16 rounding of the activations averages out least: 5e-2 here; the 2e-2 bar of north_star is held
    # at BASELINE's real sizes (tests/test_gpu_parity_full.py: B = 1024 and 4096, same criterion)
    rows = []
    for (name, po), (_, pn) in zip(o.named_parameters(), n.named_parameters()):
        go, gn = po.grad.double(), pn.grad.double().cpu()
        scale = float(go.abs().max())
        rows.append((float((go - gn).abs().max()) / (scale + 1e-30), name))
    rows.sort(reverse=True)
    bad = [r for r in rows if r[0] > 5e-2]
    assert not bad, f"gradient tensors beyond 5e-2 of their scale (err/scale, name): {bad[:8]}"


def test_ppo_minibatch_step_parity(cuda_lib):
    from multimodal_sc_b200.ppo import PPOAgent
    o, n = _pair()
    B = 8
    cam, bev = _inputs(B, seed=2)
    g = torch.Generator().manual_seed(4)
    act = torch.randint(0, SPEC.n_actions, (B,), generator=g)
    lp_old = torch.full((B,), -2.2) + 0.1 * torch.randn(B, generator=g)
    adv, ret = torch.randn(B, generator=g), torch.randn(B, generator=g)
    ao = OracleAgent(o, SPEC, minibatch=B, seed=1234)
    an = PPOAgent(n, SPEC, minibatch=B, seed=1234)
    p_before = {k: v.clone() for k, v in o.state_dict().items()}
    so = ao.minibatch_step(cam, bev, act, lp_old, adv, None, ret)
    sn = an.minibatch_step(cam.cuda(), bev.cuda(), act.cuda(), lp_old.cuda(), adv.cuda(), None, ret.cuda())
    for k in ("pi", "v", "ent", "kl", "l_sem", "total"):
        assert abs(sn[k] - so[k]) <= 2e-2 * max(abs(so[k]), 0.05), (k, sn[k], so[k])
    assert sn["clipfrac"] == pytest.approx(so["clipfrac"], abs=1.0 / B + 1e-6)
    # first Adam step = lr * g / (|g| + eps), i.e. ~lr * sign(g): wherever the oracle moved a weight by ~lr the native update must
    # have the same sign.  The strict form (update within 2e-2 lr wherever the gradient is resolved) runs at B = 4096 in
    # tests/test_gpu_parity_full.py::test_cfg3_minibatch_step_b4096
    agree, total = 0, 0
    for k, v in n.state_dict().items():
        do = (o.state_dict()[k] - p_before[k]).flatten()
        dn = (v.cpu() - p_before[k]).flatten()
        m = do.abs() > 0.9 * SPEC.lr
        agree += int((torch.sign(do[m]) == torch.sign(dn[m])).sum())
        total += int(m.sum())
    assert agree / total > 0.97, agree / total


def test_update_runs_and_learns_value(cuda_lib):
    """full update() over a tiny rollout: stats finite, value loss decreases over repeated updates."""
    from multimodal_sc_b200.model import SemComPolicy
    from multimodal_sc_b200.ppo import PPOAgent
    T, N = 8, 4
    g = torch.Generator().manual_seed(0)
    ro = dict(cam=torch.rand(T, N, 3, 128, 128, generator=g), bev=torch.rand(T, N, 1, 64, 64, generator=g),
              actions=torch.randint(0, 9, (T, N), generator=g), logp=torch.full((T, N), -2.2),
              values=torch.zeros(T, N), rewards=torch.ones(T, N), dones=torch.zeros(T, N),
              last_value=torch.zeros(N))
    pol = SemComPolicy(SPEC)
    ag = PPOAgent(pol, SPEC, minibatch=16, epochs=2)
    s0 = ag.update(ro)
    for _ in range(3):
        s1 = ag.update(ro)
    assert all(v == v and abs(v) < 1e6 for v in s1.values()), s1
    assert s1["v"] < s0["v"], (s0["v"], s1["v"])


def test_dp_shards_equal_full_batch_on_gpu(cuda_lib):
    """§8e on the native path: the sum of two 'rank' shards' gradients (global noise counters via
    sample_offset, per-sample terms scaled by 1/global_batch, L_sem by 1/world) equals the gradient of
    the full minibatch — the two ranks are emulated sequentially on one GPU through the same engine."""
    from multimodal_sc_b200 import dp, ops
    from multimodal_sc_b200.model import SemComPolicy
    pol = SemComPolicy(SPEC, init_std=0.05, seed=2)
    eng, flat = pol.engine, pol.flat
    B = 8
    cam, bev = _inputs(B, seed=3)
    cam, bev = cam.cuda(), bev.cuda()
    g = torch.Generator().manual_seed(1)
    act = torch.randint(0, SPEC.n_actions, (B,), generator=g).cuda()
    lp = (torch.full((B,), -2.2) + 0.1 * torch.randn(B, generator=g)).cuda()
    adv, ret = torch.randn(B, generator=g).cuda(), torch.randn(B, generator=g).cuda()

    def run(lo, hi, world):
        n = hi - lo
        inv_m, sem_coef, _ = dp.loss_scales(n, world, SPEC.c_sem)
        stats = torch.zeros(8, device="cuda")
        ho, A = eng.forward(cam[lo:hi].contiguous(), bev[lo:hi].contiguous(), 8.0, "rayleigh", 99, None, None, lo,
                            train=True)
        eng.l_sem(A, n, stats[5:6], grad_coef=sem_coef)
        dho = torch.empty_like(ho)
        ops.ppo_loss(ho, act[lo:hi].contiguous(), lp[lo:hi].contiguous(), adv[lo:hi].contiguous(),
                     ret[lo:hi].contiguous(), SPEC.n_actions, SPEC.clip_eps, SPEC.c_v, SPEC.c_ent, inv_m, dho=dho,
                     stats=stats)
        eng.backward(A, n, dho, None, "rayleigh", 8.0, 99, None, lo)
        return stats

    flat.g32.zero_()
    s_full = run(0, B, 1)
    g_full = flat.g32.clone()
    flat.g32.zero_()
    s_a = run(0, B // 2, 2)
    s_b = run(B // 2, B, 2)
    g_dp = flat.g32.clone()
    cos = float((g_full.double() @ g_dp.double()) / (g_full.double().norm() * g_dp.double().norm()))
    assert cos > 0.9999, cos
    assert abs(float(g_dp.norm() / g_full.norm()) - 1.0) < 2e-3
    s_sum = s_a + s_b
    s_sum[5] = s_sum[5] / 2
    assert torch.allclose(s_sum[:6].cpu(), s_full[:6].cpu(), rtol=2e-3, atol=1e-5), (s_sum, s_full)


def test_forward_single_sample_and_large_batch_properties(cuda_lib):
    """edge sizes: B = 1 runs; at B = 2048 the channel keeps unit power and the requested SNR."""
    from multimodal_sc_b200.model import SemComPolicy
    from oracle import channel as och
    pol = SemComPolicy(SPEC, init_std=0.05, seed=4)
    cam, bev = _inputs(1, seed=7)
    with torch.no_grad():
        l, v, aux = pol(cam.cuda(), bev.cuda(), 10.0, seed=3)
    assert l.shape == (1, SPEC.n_actions) and torch.isfinite(l).all() and torch.isfinite(v).all()
    B = 2048
    g = torch.Generator(device="cuda").manual_seed(0)
    camb = torch.rand(B, 3, 128, 128, generator=g, device="cuda", dtype=torch.float32).bfloat16()
    bevb = torch.rand(B, 1, 64, 64, generator=g, device="cuda", dtype=torch.float32).bfloat16()
    with torch.no_grad():
        l, v, aux = pol(camb, bevb, 5.0, channel="awgn", seed=11)
    assert torch.isfinite(l).all()
    assert abs(float(aux["z"].pow(2).mean()) - 1.0) < 1e-3
    assert abs(och.empirical_snr_db(aux["z"].cpu(), aux["y"].cpu()) - 5.0) < 0.1


def test_forward_matches_committed_golden_vectors(cuda_lib):
    """native forward at the default init / seed 0 / B = 2 against tests/golden/oracle_v0.pt (the frozen
    oracle outputs of scripts/make_golden.py), bf16 tolerance."""
    import os
    from multimodal_sc_b200 import synthetic
    from multimodal_sc_b200.model import SemComPolicy
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_v0.pt"))
    pol = SemComPolicy(SPEC)                       # same init rule and host RNG stream as the oracle
    mb = synthetic.minibatch(SPEC, 2, seed=0)
    with torch.no_grad():
        for kind in ("awgn", "rayleigh"):
            l, v, aux = pol(mb["cam"].cuda(), mb["bev"].cuda(), 10.0, channel=kind, seed=1234)
            assert _rel(l, g[f"fwd_{kind}_logits"]) < 2e-2, (kind, "logits", _rel(l, g[f"fwd_{kind}_logits"]))
            assert _rel(v, g[f"fwd_{kind}_value"]) < 2e-2, (kind, "value")
            assert abs(float(aux["l_sem"]) - float(g[f"fwd_{kind}_l_sem"])) < 2e-2 * float(g[f"fwd_{kind}_l_sem"])
            assert _rel(aux["y"][:, :, :8], g[f"fwd_{kind}_y_head"]) < 2e-2
            assert _rel(aux["z"][:, :, :8], g[f"fwd_{kind}_z_head"]) < 2e-2
        g1 = torch.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_v1.pt"))      # round-1 additions
        l, v, aux = pol(mb["cam"].cuda(), mb["bev"].cuda(), 10.0, channel="rician", seed=1234)
        assert _rel(l, g1["fwd_rician_logits"]) < 2e-2 and _rel(v, g1["fwd_rician_value"]) < 2e-2
        assert abs(float(aux["l_sem"]) - float(g1["fwd_rician_l_sem"])) < 2e-2 * float(g1["fwd_rician_l_sem"])
        assert _rel(aux["y"][:, :, :8], g1["fwd_rician_y_head"]) < 2e-2


@pytest.mark.parametrize("name,switches", [
    ("single-CTA wgrads", {"cta2_wgrad": 0}),
    ("single-CTA long-K dgrads", {"cta2_mink": 0}),
    ("two staging slots / three stages for the output-only short-K epilogues", {"one_slot": 0}),
    ("one-warpgroup fused LayerNorm", {"ln_split": 0}),
    ("split fused LayerNorm on every K", {"ln_split": 2}),
    ("fused-MLP backward without weight multicast", {"mlp_dbg": 4}),
    ("programmatic dependent launch", {"pdl": 1}),
    ("alternating traversal direction", {"alternate": 1}),
    ("L2 prefetch switches", {"gemm_dbg": 256, "mlp_dbg": 3}),
    ("wgrads on a side stream", {"side_stream": 1}),
    ("weight tiles shared by CTA clusters (multicast)", {"mcb": 1}),
    ("column-split epilogue for one-output GEMMs", {"col_split": 1}),
    ("column-split epilogue also for two outputs", {"col_split": 2}),
    ("both encoders on one stream (the default at this size is two)", {"single_encoder_stream": 1}),
])
def test_schedule_switches_preserve_results(cuda_lib, name, switches):
    """Every schedule / kernel-variant switch of the library (profiles/README.md, negative-results table) computes the
    same PPO step as the default configuration: loss statistics and the flat gradient bucket agree to bf16 level.  B = 256 so that the encoder GEMMs have 128 row tiles (pairs, multicast and both
    traversal directions are exercised on more than one tile per CTA)."""
    from multimodal_sc_b200 import ops, synthetic
    from multimodal_sc_b200.model import SemComPolicy
    from multimodal_sc_b200.ppo import PPOAgent
    B = 256
    mb = synthetic.minibatch(SPEC, B, seed=21, device="cuda")
    defaults = {"cta2_wgrad": 1, "ln_split": 1, "mlp_dbg": 0, "pdl": 0, "alternate": 0, "gemm_dbg": 0, "mcb": 0, "col_split": 0,
                "cta2_mink": 512, "one_slot": 1}

    def run(sw):
        for k, v in {**defaults, **{k: v for k, v in sw.items() if k in defaults}}.items():
            ops.debug_set(k, v)
        pol = SemComPolicy(SPEC, seed=0)
        if sw.get("side_stream"):
            pol.engine.wgrad_stream = torch.cuda.Stream()
        if sw.get("single_encoder_stream"):
            pol.engine.enc_overlap = False
        else:
            assert pol.engine._enc_side_stream(B) is not None or sw.get("side_stream")
        ag = PPOAgent(pol, SPEC, minibatch=B, seed=5)
        grads = {}
        orig = ops.adam_flat

        def spy(p32, g32, *a, **k):
            grads["g"] = g32.clone()
            return orig(p32, g32, *a, **k)

        ops.adam_flat = spy
        try:
            st = ag.minibatch_step(mb["cam"], mb["bev"], mb["actions"], mb["logp_old"], mb["adv"], None, mb["ret"])
        finally:
            ops.adam_flat = orig
            for k, v in defaults.items():
                ops.debug_set(k, v)
        return st, grads["g"]

    st0, g0 = run({})
    st1, g1 = run(switches)
    # the LayerNorm variants add the row statistics in a different order: an rstd one ulp off flips individual bf16
    # roundings downstream, so the yardstick is bf16-level agreement, far below the damage a protocol bug does
    for k in st0:
        assert abs(st0[k] - st1[k]) <= 1e-4 + 2e-3 * abs(st0[k]), (name, k, st0[k], st1[k])
    scale = float(g0.abs().max())
    assert float((g0 - g1).abs().max()) <= 2e-2 * scale, (name, float((g0 - g1).abs().max()), scale)
    assert float((g0 - g1).abs().mean()) <= 5e-4 * scale, (name, float((g0 - g1).abs().mean()), scale)
