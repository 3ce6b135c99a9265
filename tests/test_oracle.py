"""CPU pins for the oracle: Philox KATs, channel invariants, closed-form GAE, hand-computed PPO loss,
fp64 gradcheck, golden vectors.  Upstream has no tests to mirror
(/root/reference/README.md:1-2 only) — these ARE the pins, and the oracle is 'parity unpinned'
against upstream."""
import math
import os

import numpy as np
import pytest
import torch

from multimodal_sc_b200 import synthetic
from multimodal_sc_b200.spec import SpecV0, fwd_flops_per_sample
from oracle import channel as och
from oracle import philox
from oracle import ppo as oppo
from oracle.model import SemComPolicy, patchify, sample_action

SPEC = SpecV0()
GOLD = os.path.join(os.path.dirname(__file__), "golden", "oracle_v0.pt")


def test_philox_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), "408f276d 41c83b0e a20bc7c6 6d5451fd"),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            "d16cfe09 94fdcceb 5001e420 24126ea1")]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(np.array(ctr, dtype=np.uint32), key)
        assert " ".join(f"{int(x):08x}" for x in got) == want


def test_uniform_is_exact_open_interval():
    x = np.array([0, 1, 511, 512, 0xffffffff], dtype=np.uint32)
    u = philox.u32_to_uniform(x)
    assert (u > 0).all() and (u < 1).all()
    assert np.array_equal(u.astype(np.float32).astype(np.float64), u)       # exactly representable in fp32


def test_noise_offset_consistency_and_moments():
    a = philox.noise_normal(5, 6, 1024, sample_offset=0)
    b = philox.noise_normal(5, 2, 1024, sample_offset=4)
    assert np.array_equal(a[4:], b)                                          # global-index counters
    n = philox.noise_normal(9, 256, 1024)
    assert abs(n.mean()) < 5e-3 and abs(n.std() - 1) < 5e-3
    h = philox.fading_coeff(9, 20000)
    assert abs((h ** 2).sum(-1).mean() - 1.0) < 2e-2                         # E|h|^2 = 1


def test_channel_invariants():
    z = torch.randn(64, 2, SPEC.k_sym) * 5
    zn = och.power_normalize(z)
    assert torch.allclose(zn.pow(2).mean(-1), torch.ones(64, 2), atol=1e-5)
    for snr in (-5.0, 0.0, 10.0, 20.0):
        y, zn = och.channel_forward(z, snr, "awgn", seed=3)
        assert abs(och.empirical_snr_db(zn, y) - snr) < 0.1
    y, zn = och.channel_forward(z, 300.0, "rayleigh", seed=3)               # no noise: MMSE inverts h exactly
    assert torch.allclose(y, zn, atol=1e-4)
    with pytest.raises(ValueError):
        och.channel_forward(z, 10.0, "awgn")
    y, _ = och.channel_forward(z, 10.0, "none")
    assert torch.equal(y, och.power_normalize(z))


def test_rician_channel_pins():
    """h = sqrt(K/(K+1)) + sqrt(1/(K+1)) CN(0,1).  E|h|^2 = 1 for every K, the mean of h is the
    line-of-sight term, K = 0 is the Rayleigh draw itself, and with no noise the MMSE receiver inverts h exactly."""
    h0 = och.make_fading(11, 200000)
    assert torch.equal(och.rician_fading(h0, 0.0), h0)
    for K in (0.5, 4.0, 30.0):
        h = och.rician_fading(h0, K)
        assert abs(float(h.pow(2).sum(-1).mean()) - 1.0) < 5e-3
        assert abs(float(h[..., 0].mean()) - (K / (K + 1)) ** 0.5) < 5e-3 and abs(float(h[..., 1].mean())) < 5e-3
        assert abs(float(h[..., 1].var()) - 0.5 / (K + 1)) < 5e-3
    z = torch.randn(16, 2, SPEC.k_sym)
    y, zn = och.channel_forward(z, 300.0, "rician", seed=3, rician_k=4.0)
    assert torch.allclose(y, zn, atol=1e-4)
    y0, _ = och.channel_forward(z, 7.0, "rician", seed=3, rician_k=0.0)
    yr, _ = och.channel_forward(z, 7.0, "rayleigh", seed=3)
    assert torch.equal(y0, yr)
    with pytest.raises(ValueError):
        och.channel_forward(z, 10.0, "rician")


@pytest.mark.parametrize("kind", ["awgn", "rayleigh", "rician"])
def test_channel_gradcheck_fp64(kind):
    z = torch.randn(2, 2, 8, dtype=torch.float64, requires_grad=True)
    noise = torch.randn(2, 2, 8, dtype=torch.float64)
    h = torch.randn(2, 2, 2, dtype=torch.float64)
    f = lambda zz: och.channel_forward(zz, torch.tensor([3.0, 12.0]), kind, noise=noise, fading=h)[0]
    assert torch.autograd.gradcheck(f, (z,), eps=1e-6, atol=1e-6)


def test_gae_closed_form_and_done_mask():
    T, N = 32, 2
    r = torch.ones(T, N)
    z = torch.zeros(T, N)
    adv, ret = oppo.gae(r, z, z, torch.zeros(N), SPEC.gamma, SPEC.lam)
    gl = SPEC.gamma * SPEC.lam
    ref = torch.tensor([(1 - gl ** (T - t)) / (1 - gl) for t in range(T)])
    assert torch.allclose(adv[:, 0], ref, rtol=1e-5)
    assert torch.equal(ret, adv)
    d = z.clone()
    d[9] = 1.0                                                               # episode ends at t = 9
    adv2, _ = oppo.gae(r, z, d, torch.full((N,), 100.0), SPEC.gamma, SPEC.lam)
    ref2 = torch.tensor([(1 - gl ** (10 - t)) / (1 - gl) for t in range(10)])
    assert torch.allclose(adv2[:10, 0], ref2, rtol=1e-5)                     # nothing leaks across the boundary
    n = oppo.normalize_adv(adv2)
    assert abs(float(n.mean())) < 1e-5 and abs(float(n.std(unbiased=False)) - 1) < 1e-4


def test_ppo_loss_hand_computed():
    A = SPEC.n_actions
    logits = torch.zeros(4, A)                       # uniform policy: logp = -ln 9, entropy = ln 9
    act = torch.tensor([0, 1, 2, 3])
    lnA = math.log(A)
    ratio = torch.tensor([1.0, 1.5, 0.5, 1.1])
    logp_old = -lnA - torch.log(ratio)
    adv = torch.tensor([1.0, 2.0, -1.0, -2.0])
    value, ret = torch.tensor([0.0, 1.0, 2.0, 3.0]), torch.tensor([1.0, 1.0, 1.0, 1.0])
    out = oppo.ppo_loss(logits, act, logp_old, adv, value, ret, torch.tensor(0.5), SPEC)
    # min(r A, clip(r) A): [1, min(3, 2.4)=2.4, min(-.5, -.8)=-.8, min(-2.2, -2.2)=-2.2]
    pi = -(1.0 + 2.4 - 0.8 - 2.2) / 4
    v = 0.5 * (1 + 0 + 1 + 4) / 4
    assert float(out["pi"]) == pytest.approx(pi, rel=1e-5)
    assert float(out["v"]) == pytest.approx(v, rel=1e-6)
    assert float(out["ent"]) == pytest.approx(lnA, rel=1e-6)
    assert float(out["clipfrac"]) == pytest.approx(0.5)
    total = pi + SPEC.c_v * v - SPEC.c_ent * lnA + SPEC.c_sem * 0.5
    assert float(out["total"]) == pytest.approx(total, rel=1e-5)


def test_sample_action_inverse_cdf():
    logits = torch.log(torch.tensor([[0.1, 0.2, 0.3, 0.4]]))
    for u, want in [(0.0, 0), (0.0999, 0), (0.1001, 1), (0.2999, 1), (0.3001, 2), (0.61, 3), (0.999999, 3)]:
        a, lp = sample_action(logits, torch.tensor([u]))
        assert int(a) == want
        assert float(lp) == pytest.approx(math.log([0.1, 0.2, 0.3, 0.4][want]), rel=1e-5)


def test_patchify_layout_and_model_shapes():
    img = torch.arange(2 * 3 * 32 * 32, dtype=torch.float32).view(2, 3, 32, 32)
    p = patchify(img, 16)
    assert p.shape == (2, 4, 768)
    assert float(p[1, 3, 256 + 16 * 2 + 5]) == float(img[1, 1, 16 + 2, 16 + 5])   # (ty=1,tx=1), c=1, py=2, px=5
    pol = SemComPolicy(SPEC)
    assert sum(x.numel() for x in pol.parameters()) == 5165866
    assert abs(fwd_flops_per_sample(SPEC) / 1e6 - 901.9) < 0.1
    mb = synthetic.minibatch(SPEC, 2, seed=0)
    l, v, aux = pol(mb["cam"], mb["bev"], 10.0, seed=1)
    assert l.shape == (2, 9) and v.shape == (2,) and aux["z"].shape == (2, 2, 1024)
    l2, _, _ = pol(mb["cam"], mb["bev"], 10.0, seed=1)
    assert torch.equal(l, l2)                                                  # deterministic under a fixed seed
    mp = oppo.minibatch_permutation(10, torch.Generator().manual_seed(3))
    assert torch.equal(mp, torch.randperm(10, generator=torch.Generator().manual_seed(3)))


def test_golden_vectors():
    g = torch.load(GOLD)
    assert np.array_equal(g["philox_words_seed1234_off5"].numpy().astype(np.uint32), philox.philox_words(1234, 8, 0, 5))
    assert torch.equal(g["noise_seed7_b2_off3"], torch.from_numpy(philox.noise_normal(7, 2, 16, 3)))
    assert torch.equal(g["fading_seed7_b3_off1"], torch.from_numpy(philox.fading_coeff(7, 3, 1)))
    torch.manual_seed(0)
    pol = SemComPolicy(SPEC)
    mb = synthetic.minibatch(SPEC, 2, seed=0)
    with torch.no_grad():
        for kind in ("awgn", "rayleigh"):
            l, v, aux = pol(mb["cam"], mb["bev"], 10.0, channel=kind, seed=1234)
            assert torch.allclose(l, g[f"fwd_{kind}_logits"], rtol=1e-4, atol=1e-6), kind
            assert torch.allclose(v, g[f"fwd_{kind}_value"], rtol=1e-4, atol=1e-6)
            assert torch.allclose(aux["l_sem"], g[f"fwd_{kind}_l_sem"], rtol=1e-4)
            assert torch.allclose(aux["y"][:, :, :8], g[f"fwd_{kind}_y_head"], rtol=1e-4, atol=1e-5)
            assert torch.allclose(aux["z"][:, :, :8], g[f"fwd_{kind}_z_head"], rtol=1e-4, atol=1e-5)
    gg = torch.Generator().manual_seed(11)
    r, v = torch.randn(16, 3, generator=gg), torch.randn(16, 3, generator=gg)
    d = (torch.rand(16, 3, generator=gg) < 0.2).float()
    lv = torch.randn(3, generator=gg)
    adv, ret = oppo.gae(r, v, d, lv, SPEC.gamma, SPEC.lam)
    assert torch.allclose(adv, g["gae_adv"], rtol=1e-6) and torch.allclose(ret, g["gae_ret"], rtol=1e-6)
    assert torch.allclose(oppo.normalize_adv(adv, SPEC.adv_eps), g["gae_adv_norm"], rtol=1e-5, atol=1e-6)


def test_golden_vectors_v1_rician():
    """oracle against tests/golden/oracle_v1.pt (scripts/make_golden_v1.py): pins the Rician channel added in round 1."""
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_v1.pt"))
    gen = torch.Generator().manual_seed(11)
    z = torch.randn(3, 2, SPEC.k_sym, generator=gen) * 0.4
    snr = torch.tensor([-3.0, 6.0, 15.0])
    y, _ = och.channel_forward(z, snr, "rician", seed=99, sample_offset=5, rician_k=2.5)
    assert torch.allclose(y, g["rician_k2.5_seed99_off5_y"], rtol=1e-6, atol=1e-7)
    assert torch.allclose(och.rician_fading(och.make_fading(99, 3, 5), 2.5), g["rician_fading_seed99_b3_off5_k2.5"],
                          rtol=1e-6, atol=1e-7)


def test_golden_vectors_v2_options():
    """oracle against tests/golden/oracle_v2.pt (scripts/make_golden_v2.py): pins the round-2 spec options — b-bit quantised
    symbols with the clipped straight-through gradient, SNR-conditioned encoders, exact-erf GELU."""
    import dataclasses
    import importlib.util
    spec_ = importlib.util.spec_from_file_location("make_golden_v2", os.path.join(os.path.dirname(__file__), "..", "scripts", "make_golden_v2.py"))
    mg = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(mg)
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_v2.pt"))
    sp = SpecV0()
    gen = torch.Generator().manual_seed(23)
    z = torch.randn(3, 2, sp.k_sym, generator=gen) * 0.4
    z[0, 0, :4] *= 30.0
    snr = torch.tensor([0.0, 8.0, 16.0])
    dy = torch.randn(3, 2, sp.k_sym, generator=gen)
    for bits in (2, 5):
        zr = z.clone().requires_grad_(True)
        y, zn = och.channel_forward(zr, snr, "awgn", seed=7, quant_bits=bits, quant_clip=2.0)
        (y * dy).sum().backward()
        assert torch.allclose(y, g[f"quant{bits}_clip2_awgn_seed7_y"], rtol=1e-6, atol=1e-6)
        assert torch.allclose(zr.grad, g[f"quant{bits}_clip2_awgn_seed7_dz"], rtol=1e-5, atol=1e-6)
        # properties: the noiseless transmit symbols lie on the 2^bits-level grid; the clipped symbols get no gradient
        tx = och.quantize_ste(zn.detach(), bits, 2.0)
        step = 4.0 / 2 ** bits
        lv = tx / step - 0.5
        assert float((lv - lv.round()).abs().max()) < 1e-5 and int(torch.unique(lv.round()).numel()) <= 2 ** bits
        assert float(tx.abs().max()) <= 2.0
    mb = synthetic.minibatch(sp, 2, seed=0)
    snr2 = torch.tensor([2.0, 14.0])
    base = None
    for name, kw in mg.OPTIONS.items():
        pol = mg.policy_for(dataclasses.replace(sp, **kw))
        with torch.no_grad():
            l, v, aux = pol(mb["cam"], mb["bev"], snr2, channel="rayleigh", seed=1234)
        assert torch.allclose(l, g[f"fwd_{name}_logits"], rtol=1e-4, atol=1e-5), name
        assert torch.allclose(v, g[f"fwd_{name}_value"], rtol=1e-4, atol=1e-5), name
        assert torch.allclose(aux["l_sem"], g[f"fwd_{name}_l_sem"], rtol=1e-4), name
        assert torch.allclose(aux["y"][:, :, :8], g[f"fwd_{name}_y_head"], rtol=1e-4, atol=1e-5), name
    # the options change the outputs (the vectors are not accidentally those of the default spec)
    assert not torch.allclose(g["fwd_snr_cond_logits"], g["fwd_erf_logits"], rtol=1e-3, atol=1e-4)
