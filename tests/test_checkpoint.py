"""Checkpoint format: round trip through the oracle agent on the CPU, oracle <-> native on the GPU."""
import io

import pytest
import torch

from multimodal_sc_b200 import checkpoint, synthetic
from multimodal_sc_b200.spec import SpecV0
from oracle.model import SemComPolicy as OraclePolicy
from oracle.ppo import PPOAgent as OracleAgent

SPEC = SpecV0()


def _oracle_agent(seed):
    pol = OraclePolicy(SPEC)
    pol.reset_parameters(std=0.05, seed=seed)
    return OracleAgent(pol, SPEC, minibatch=4, seed=7)


def _step(agent, mb):
    return agent.minibatch_step(mb["cam"], mb["bev"], mb["actions"], mb["logp_old"], mb["adv"], None, mb["ret"])


def test_oracle_round_trip_resumes_bit_exactly():
    torch.set_num_threads(4)
    mbs = [synthetic.minibatch(SPEC, 4, seed=30 + i) for i in range(3)]
    a = _oracle_agent(1)
    _step(a, mbs[0])
    buf = io.BytesIO()
    checkpoint.save(buf, a)
    ref = [_step(a, mbs[1]), _step(a, mbs[2])]
    perm_ref = torch.randperm(16, generator=a.gen)
    b = _oracle_agent(2)                                    # different weights: everything must come from the file
    buf.seek(0)
    st = checkpoint.load(buf, b)
    assert st["format"] == checkpoint.FORMAT and st["adam"]["step"] == 1 and b.step_count == 1
    got = [_step(b, mbs[1]), _step(b, mbs[2])]
    for r, g in zip(ref, got):
        for k in r:
            assert r[k] == g[k], (k, r[k], g[k])            # same arithmetic, same state -> identical floats
    assert torch.equal(torch.randperm(16, generator=b.gen), perm_ref)


def test_rejects_foreign_and_mismatched_checkpoints():
    a = _oracle_agent(1)
    st = checkpoint.to_state(a)
    with pytest.raises(ValueError):
        checkpoint.from_state(a, {**st, "format": "something else"})
    import dataclasses
    other = dataclasses.replace(SPEC, lr=1e-3)
    pol = OraclePolicy(other)
    with pytest.raises(ValueError):
        checkpoint.from_state(OracleAgent(pol, other, minibatch=4), st)
    bad = {**st, "model": {k: v for k, v in list(st["model"].items())[1:]}}
    with pytest.raises(ValueError):
        checkpoint.from_state(a, bad)


@pytest.mark.gpu
def test_oracle_checkpoint_resumes_on_the_native_agent_and_back(cuda_lib):
    """two steps on the oracle -> checkpoint -> native agent: its next step matches the oracle's next step (bf16
    tolerance), and a checkpoint written by the native agent loads back into an oracle agent with the same state."""
    from multimodal_sc_b200.model import SemComPolicy
    from multimodal_sc_b200.ppo import PPOAgent
    B = 8
    mbs = [synthetic.minibatch(SPEC, B, seed=40 + i) for i in range(3)]
    o = _oracle_agent(3)
    o.minibatch = B
    _step(o, mbs[0])
    _step(o, mbs[1])
    st = checkpoint.to_state(o)
    n = PPOAgent(SemComPolicy(SPEC, seed=9), SPEC, minibatch=B, seed=7)
    checkpoint.from_state(n, st)
    assert n.step_count == 2 and int(n.dev_step) == 2
    s_o = _step(o, mbs[2])
    s_n = _step(n, {k: v.cuda() for k, v in mbs[2].items()})
    for k in ("pi", "v", "ent", "l_sem"):
        assert abs(s_n[k] - s_o[k]) <= 2e-2 * max(abs(s_o[k]), 0.05), (k, s_n[k], s_o[k])
    # native -> oracle: moments and weights survive the trip
    st2 = checkpoint.to_state(n)
    o2 = _oracle_agent(4)
    checkpoint.from_state(o2, st2)
    assert o2.step_count == 3
    for name, p in o2.policy.named_parameters():
        assert torch.allclose(p.detach(), st2["model"][name])
        assert torch.allclose(o2.opt.state[p]["exp_avg"], st2["adam"]["m"][name])
    # after one more (identical) update the two implementations hold nearly the same Adam moments
    rel = []
    for name, p in o.policy.named_parameters():
        mo, mn = o.opt.state[p]["exp_avg"], st2["adam"]["m"][name]
        rel.append(float((mo - mn).norm() / (mo.norm() + 1e-12)))
    assert max(rel) < 0.15 and sum(rel) / len(rel) < 0.05, (max(rel), sum(rel) / len(rel))
