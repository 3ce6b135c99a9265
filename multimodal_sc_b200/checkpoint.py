"""Checkpoints of a PPO agent, interchangeable between the native agent and the oracle.

One dict, saved with `torch.save`: the policy's `state_dict` (same keys and shapes on both sides), the Adam moments
under the same parameter names, the optimizer step, the minibatch-permutation generator state and the agent's step
counter (which also seeds the channel noise).  A checkpoint written by the oracle's `PPOAgent` (torch.optim.Adam)
loads into the native `PPOAgent` (flat fp32 moment buffers updated by `sc_adam_flat`) and vice versa, so training can
move between the two implementations mid-run — which is also how the format is tested.  Upstream has no checkpoint
format to follow (/root/reference/README.md:1-2); if it publishes one, only `to_state` / `from_state` change.
"""
from __future__ import annotations

import dataclasses

import torch

FORMAT = "semcom-ppo-checkpoint-v1"


def _is_native(agent) -> bool:
    return hasattr(agent.policy, "flat")


def to_state(agent) -> dict:
    """agent (native or oracle PPOAgent) -> checkpoint dict (CPU tensors only)."""
    pol = agent.policy
    model = {k: v.detach().to("cpu", torch.float32).clone() for k, v in pol.state_dict().items()}
    if _is_native(agent):
        flat = pol.flat
        # data parallel with the peer-memory optimizer step: every rank maintains the Adam moments of its own slice only, so
        # they are completed first — a COLLECTIVE: every rank of the group must call to_state() / save() (any of them may
        # then write the file; the result is identical on all ranks)
        if getattr(agent, "_fused", None) is not None:
            agent._fused.gather_moments()
        m = {n: flat._view(flat.m, n).detach().to("cpu").clone() for n in flat.shapes}
        v = {n: flat._view(flat.v, n).detach().to("cpu").clone() for n in flat.shapes}
        opt_step = int(agent.step_count)
    else:
        named = dict(pol.named_parameters())
        st = agent.opt.state
        m = {n: (st[p]["exp_avg"].detach().cpu().clone() if p in st else torch.zeros_like(p, device="cpu"))
             for n, p in named.items()}
        v = {n: (st[p]["exp_avg_sq"].detach().cpu().clone() if p in st else torch.zeros_like(p, device="cpu"))
             for n, p in named.items()}
        steps = {int(st[p]["step"]) for p in named.values() if p in st}
        assert len(steps) <= 1, f"parameters disagree on the optimizer step: {steps}"
        opt_step = steps.pop() if steps else 0
    return {"format": FORMAT, "spec": dataclasses.asdict(agent.spec), "model": model,
            "adam": {"m": m, "v": v, "step": opt_step},
            "step_count": int(agent.step_count), "perm_rng": agent.gen.get_state().clone(),
            "config": {"epochs": agent.epochs, "minibatch": agent.minibatch, "snr_db": agent.snr_db,
                       "channel": agent.channel, "seed": agent.seed}}


def from_state(agent, state: dict, *, strict_spec: bool = True) -> None:
    """load a checkpoint dict into `agent` (native or oracle), in place."""
    if state.get("format") != FORMAT:
        raise ValueError(f"not a {FORMAT} checkpoint: format={state.get('format')!r}")
    if strict_spec and state["spec"] != dataclasses.asdict(agent.spec):
        diff = {k: (state["spec"].get(k), v) for k, v in dataclasses.asdict(agent.spec).items() if state["spec"].get(k) != v}
        raise ValueError(f"checkpoint was written for a different SpecV0: {diff}")
    pol = agent.policy
    missing = set(pol.state_dict()) ^ set(state["model"])
    if missing:
        raise ValueError(f"parameter names differ: {sorted(missing)[:5]} ...")
    pol.load_state_dict(state["model"])
    adam = state["adam"]
    if _is_native(agent):
        flat = pol.flat
        for n in flat.shapes:
            flat._view(flat.m, n).copy_(adam["m"][n])
            flat._view(flat.v, n).copy_(adam["v"][n])
        if adam["step"] != state["step_count"]:
            raise ValueError("native agent keeps one counter: optimizer step and step_count must agree")
        agent.step_count = int(state["step_count"])
        agent.dev_step.fill_(agent.step_count)
    else:
        agent.opt.state.clear()
        for n, p in pol.named_parameters():
            if adam["step"] > 0:
                agent.opt.state[p] = {"step": torch.tensor(float(adam["step"])),
                                      "exp_avg": adam["m"][n].to(p.device, p.dtype).clone(),
                                      "exp_avg_sq": adam["v"][n].to(p.device, p.dtype).clone()}
        agent.step_count = int(state["step_count"])
    agent.gen.set_state(state["perm_rng"])


def save(path, agent) -> None:
    torch.save(to_state(agent), path)


def load(path, agent, **kw) -> dict:
    state = torch.load(path, map_location="cpu", weights_only=True)
    from_state(agent, state, **kw)
    return state
