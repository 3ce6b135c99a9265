"""Seeded synthetic inputs of the SPEC v0 shapes: camera U[0,1), sparse BEV
occupancy, N(0,1) rewards, Bernoulli(1/64) dones.  Host-side data plumbing only."""
from __future__ import annotations

import torch

from .spec import SpecV0


def minibatch(spec: SpecV0, batch: int, seed: int = 0, device="cpu", pin: bool = False,
              obs_dtype=torch.float32) -> dict:
    """one PPO minibatch: cam, bev, actions, logp_old, adv, ret (fp32 / int64).  obs_dtype=bfloat16 gives the
    'fp32 -> bf16 once, outside' form of SURVEY.md section 8d (cast on the host, before any timed region)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    cam = torch.rand(batch, spec.cam_ch, spec.cam_hw, spec.cam_hw, generator=g)
    occ = torch.rand(batch, spec.bev_ch, spec.bev_hw, spec.bev_hw, generator=g) < 0.1
    bev = occ.float() * torch.rand(batch, spec.bev_ch, spec.bev_hw, spec.bev_hw, generator=g)
    out = dict(
        cam=cam, bev=bev,
        actions=torch.randint(0, spec.n_actions, (batch,), generator=g),
        logp_old=torch.full((batch,), -2.1972246) + 0.05 * torch.randn(batch, generator=g),
        adv=torch.randn(batch, generator=g),
        ret=torch.randn(batch, generator=g),
    )
    out["cam"], out["bev"] = out["cam"].to(obs_dtype), out["bev"].to(obs_dtype)
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    if device != "cpu":
        out = {k: v.to(device) for k, v in out.items()}
    return out


def rollout(spec: SpecV0, T: int, N: int, seed: int = 0) -> dict:
    g = torch.Generator().manual_seed(seed)
    occ = torch.rand(T, N, spec.bev_ch, spec.bev_hw, spec.bev_hw, generator=g) < 0.1
    return dict(
        cam=torch.rand(T, N, spec.cam_ch, spec.cam_hw, spec.cam_hw, generator=g),
        bev=occ.float() * torch.rand(T, N, spec.bev_ch, spec.bev_hw, spec.bev_hw, generator=g),
        actions=torch.randint(0, spec.n_actions, (T, N), generator=g),
        logp=torch.full((T, N), -2.1972246),
        values=0.1 * torch.randn(T, N, generator=g),
        rewards=torch.randn(T, N, generator=g),
        dones=(torch.rand(T, N, generator=g) < 1.0 / 64).float(),
        last_value=0.1 * torch.randn(N, generator=g),
    )
