"""Rollout collection around the hot path (the caller of `act()` / `update()`).

Upstream (/root/reference/README.md:1-2) ships no environment, trainer or rollout code, so this module is SPEC-side:
  * `SyntheticDrivingEnv` — a vectorised lane-keeping / obstacle-avoidance stub that renders the two observation
    modalities of SPEC v0 (camera 3x128x128, LiDAR-BEV 1x64x64) from a 5-number state with plain torch ops.  It runs
    on the CPU (tests) or on the GPU next to the policy (no host round trip per step);
  * `GraphedAct` — `SemComPolicy.act` for a fixed number of environments captured once as a CUDA graph
    (inference forward with the fused MLP kernels + inverse-CDF sampling, ~60 launches) and replayed per step.  The
    channel's Philox key is `seed + counter` with the counter on the device, so every replay draws fresh noise;
  * `RolloutCollector` — T steps of act -> env.step into preallocated [T, N, ...] buffers, returning the dict
    `PPOAgent.update` consumes.
"""
from __future__ import annotations

import math

import torch

from .spec import SpecV0

BF16, F32 = torch.bfloat16, torch.float32


class SyntheticDrivingEnv:
    """N independent cars on a straight three-lane road.

    state per env: lateral offset x in lane widths (|x| > 1.5 = off road), heading th (rad), speed v in [0, 1],
    distance d to the obstacle ahead (in view for d in (0, 1]), obstacle lateral position o.
    action a in [0, 9): steer = a % 3 - 1, accel = a // 3 - 1.
    reward = progress - lane-centre cost - collision / off-road penalty; episodes end on collision, off-road or
    after `max_steps`, and reset themselves (seeded generator: fully reproducible)."""

    N_ACTIONS = 9

    def __init__(self, n_envs: int, device="cpu", seed: int = 0, max_steps: int = 256, spec: SpecV0 = SpecV0()):
        assert spec.n_actions == self.N_ACTIONS
        self.n, self.device, self.spec, self.max_steps = n_envs, torch.device(device), spec, max_steps
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        z = lambda: torch.zeros(n_envs, device=self.device, dtype=F32)
        self.x, self.th, self.v, self.d, self.o = z(), z(), z(), z(), z()
        self.t = torch.zeros(n_envs, device=self.device, dtype=torch.int64)
        H, W = spec.cam_hw, spec.cam_hw
        self._cam_rows = torch.linspace(0.0, 1.0, H, device=self.device).view(1, H, 1)        # 0 = horizon, 1 = bumper
        self._cam_cols = torch.linspace(-1.0, 1.0, W, device=self.device).view(1, 1, W)
        Hb = spec.bev_hw
        self._bev_rows = torch.linspace(1.0, 0.0, Hb, device=self.device).view(1, Hb, 1)       # distance ahead
        self._bev_cols = torch.linspace(-2.0, 2.0, Hb, device=self.device).view(1, 1, Hb)      # lateral, lane widths
        self._reset(torch.ones(n_envs, dtype=torch.bool, device=self.device))

    def _rand(self, n):
        return torch.rand(n, generator=self.gen, device=self.device, dtype=F32)

    def _reset(self, mask):
        r = self._rand(self.n * 4).view(4, self.n)
        self.x = torch.where(mask, (r[0] - 0.5) * 1.0, self.x)
        self.th = torch.where(mask, (r[1] - 0.5) * 0.2, self.th)
        self.v = torch.where(mask, 0.3 + 0.4 * r[2], self.v)
        self.d = torch.where(mask, torch.ones_like(self.d), self.d)
        self.o = torch.where(mask, torch.round(r[3] * 2.0 - 1.0), self.o)                    # lane -1, 0 or +1
        self.t = torch.where(mask, torch.zeros_like(self.t), self.t)

    # ---- observations -------------------------------------------------------------------------
    def observe(self, dtype=BF16):
        """(cam [N,3,128,128], bev [N,1,64,64]) in [0,1]."""
        x, th, d, o = (t.view(-1, 1, 1) for t in (self.x, self.th, self.d, self.o))
        rows, cols = self._cam_rows, self._cam_cols
        # camera: perspective road.  A lane line at lateral offset l projects to column (l - x) * rows + th * (1 - rows)
        centre = th * (1.0 - rows)
        lines = torch.zeros(self.n, rows.shape[1], cols.shape[2], device=self.device, dtype=F32)
        for l in (-1.5, -0.5, 0.5, 1.5):
            lines = torch.maximum(lines, ((cols - ((l - x) * 0.6 * rows + centre)).abs() < 0.012 + 0.02 * rows).float())
        size = 0.05 + 0.35 * (1.0 - d)                                                        # the obstacle grows as it nears
        obst = (((cols - ((o - x) * 0.6 * (1.0 - d) + centre)).abs() < size) &
                ((rows - (1.0 - d) * 0.9).abs() < size)).float()
        sky = (1.0 - rows).expand(self.n, -1, cols.shape[2])
        cam = torch.stack([sky, lines, obst], dim=1)
        # BEV occupancy: lane borders and the obstacle box, ego at the bottom centre
        br, bc = self._bev_rows, self._bev_cols
        border = ((bc - (-1.5 - x)).abs() < 0.06) | ((bc - (1.5 - x)).abs() < 0.06)
        box = ((bc - (o - x)).abs() < 0.35) & ((br - d).abs() < 0.08)
        bev = (border.expand(self.n, br.shape[1], -1) | box).float().unsqueeze(1)
        return cam.to(dtype), bev.to(dtype)

    # ---- dynamics -----------------------------------------------------------------------------
    def step(self, action):
        """action int64 [N] -> (reward fp32 [N], done fp32 [N]); finished episodes are reset in place."""
        a = action.to(self.device)
        steer = (a % 3 - 1).to(F32)
        accel = (a // 3 - 1).to(F32)
        self.th = (self.th + 0.08 * steer).clamp(-0.6, 0.6)
        self.v = (self.v + 0.05 * accel).clamp(0.0, 1.0)
        self.x = self.x + self.v * torch.sin(self.th) * 0.5
        self.d = self.d - self.v * torch.cos(self.th) * 0.04
        self.t = self.t + 1
        hit = (self.d <= 0.05) & ((self.x - self.o).abs() < 0.5)
        passed = (self.d <= 0.0) & ~hit
        off = self.x.abs() > 1.5
        timeout = self.t >= self.max_steps
        reward = self.v * torch.cos(self.th) - 0.2 * self.x.abs() - 5.0 * hit.float() - 5.0 * off.float()
        done = hit | off | timeout
        # a passed obstacle is replaced by a new one at the horizon
        r = self._rand(self.n)
        self.d = torch.where(passed, torch.ones_like(self.d), self.d)
        self.o = torch.where(passed, torch.round(r * 2.0 - 1.0), self.o)
        self._reset(done)
        return reward, done.to(F32)


class GraphedAct:
    """`policy.act` for a fixed batch of N environments as one CUDA graph replay per step."""

    def __init__(self, policy, n_envs: int, snr_db: float = 10.0, channel: str = "awgn", seed: int = 0,
                 obs_dtype=BF16, use_graph: bool = True):
        from . import ops
        self.policy, self.n, self.snr_db, self.channel, self.seed = policy, n_envs, snr_db, channel, seed
        sp, dev = policy.spec, policy.device
        self.cam = torch.zeros(n_envs, 3, sp.cam_hw, sp.cam_hw, device=dev, dtype=obs_dtype)
        self.bev = torch.zeros(n_envs, sp.bev_ch, sp.bev_hw, sp.bev_hw, device=dev, dtype=obs_dtype)
        self.u = torch.zeros(n_envs, device=dev, dtype=F32)
        self.action = torch.zeros(n_envs, device=dev, dtype=torch.int64)
        self.logp = torch.zeros(n_envs, device=dev, dtype=F32)
        self.value = torch.zeros(n_envs, device=dev, dtype=F32)
        self.counter = torch.zeros(1, device=dev, dtype=torch.int64)      # steps taken: the channel key is seed + counter
        self._ops = ops
        self.graph = None
        self._body()                                                       # warm-up: arena, kernel attributes
        self.counter.zero_()
        if use_graph:
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._body()

    def _body(self):
        pol, sp = self.policy, self.policy.spec
        ho, _ = pol.engine.forward(self.cam, self.bev, self.snr_db, self.channel, self.seed, None, None, 0,
                                   train=False, seed_dev=self.counter, need_recon=False)
        self._ops.sample_action(ho, self.u, sp.n_actions, action=self.action, logp=self.logp)
        self.value.copy_(ho[:, sp.n_actions])
        self.counter.add_(1)

    @torch.no_grad()
    def __call__(self, cam, bev, u):
        """returns (action int64 [N], logp [N], value [N]) — views of static buffers, valid until the next call."""
        self.cam.copy_(cam, non_blocking=True)
        self.bev.copy_(bev, non_blocking=True)
        self.u.copy_(u, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()
        return self.action, self.logp, self.value


class RolloutCollector:
    """T steps of act -> env.step into preallocated [T, N, ...] buffers (the dict `PPOAgent.update` takes)."""

    def __init__(self, act, env: SyntheticDrivingEnv, horizon: int, spec: SpecV0 = SpecV0(), seed: int = 0,
                 obs_dtype=BF16, device=None):
        self.act, self.env, self.T, self.spec = act, env, horizon, spec
        dev = torch.device(device) if device is not None else env.device
        N = env.n
        self.gen = torch.Generator(device=dev).manual_seed(seed)
        self.device = dev
        self.buf = {
            "cam": torch.empty(horizon, N, 3, spec.cam_hw, spec.cam_hw, device=dev, dtype=obs_dtype),
            "bev": torch.empty(horizon, N, spec.bev_ch, spec.bev_hw, spec.bev_hw, device=dev, dtype=obs_dtype),
            "actions": torch.empty(horizon, N, device=dev, dtype=torch.int64),
            "logp": torch.empty(horizon, N, device=dev, dtype=F32),
            "values": torch.empty(horizon, N, device=dev, dtype=F32),
            "rewards": torch.empty(horizon, N, device=dev, dtype=F32),
            "dones": torch.empty(horizon, N, device=dev, dtype=F32),
            "last_value": torch.empty(N, device=dev, dtype=F32),
        }
        self.obs_dtype = obs_dtype

    @torch.no_grad()
    def collect(self) -> dict:
        b, env = self.buf, self.env
        for t in range(self.T):
            cam, bev = env.observe(self.obs_dtype)
            u = torch.rand(env.n, generator=self.gen, device=self.device, dtype=F32)
            a, logp, v = self.act(cam.to(self.device), bev.to(self.device), u)
            b["cam"][t].copy_(cam)
            b["bev"][t].copy_(bev)
            b["actions"][t].copy_(a)
            b["logp"][t].copy_(logp)
            b["values"][t].copy_(v)
            r, d = env.step(a)
            b["rewards"][t].copy_(r)
            b["dones"][t].copy_(d)
        cam, bev = env.observe(self.obs_dtype)
        u = torch.rand(env.n, generator=self.gen, device=self.device, dtype=F32)
        _, _, v = self.act(cam.to(self.device), bev.to(self.device), u)        # bootstrap value of the state after step T
        b["last_value"].copy_(v)
        return b

    @staticmethod
    def steps_per_second(collector: "RolloutCollector", repeats: int = 2) -> float:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(repeats):
            collector.collect()
        e1.record()
        torch.cuda.synchronize()
        return repeats * collector.T * collector.env.n / (e0.elapsed_time(e1) * 1e-3)


def _check_obs_ranges(cam, bev):  # used by the tests
    return bool((cam.float() >= 0).all() and (cam.float() <= 1).all() and (bev.float() >= 0).all() and
                (bev.float() <= 1).all() and math.isfinite(float(cam.float().sum())))
