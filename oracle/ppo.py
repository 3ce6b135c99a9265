"""GAE, the PPO clipped loss and the update loop — eager-PyTorch restatement of SPEC v0
rows a8, a9, a10.

Reference citation: no upstream trainer exists (/root/reference/README.md:1-2 only); this
follows SURVEY.md §3.2(B) and §8 rows a8-a10 (Schulman et al. 2016/2017 for GAE/PPO).
"""
from __future__ import annotations

import torch

from multimodal_sc_b200.spec import SpecV0


def gae(rewards, values, dones, last_value, gamma: float, lam: float):
    """a8.  rewards, values, dones: [T,N]; last_value: [N].  dones[t]=1 means the episode ended
    AT step t (no bootstrap from t+1).  Returns (adv[T,N], ret[T,N]) in fp32."""
    T, N = rewards.shape
    r = rewards.float()
    v = values.float()
    nd = 1.0 - dones.float()
    adv = torch.zeros(T, N, dtype=torch.float32, device=r.device)
    nxt_v = last_value.float()
    nxt_a = torch.zeros(N, dtype=torch.float32, device=r.device)
    for t in range(T - 1, -1, -1):
        delta = r[t] + gamma * nxt_v * nd[t] - v[t]
        nxt_a = delta + gamma * lam * nd[t] * nxt_a
        adv[t] = nxt_a
        nxt_v = v[t]
    return adv, adv + v


def normalize_adv(adv: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """(A - mean) / (std + eps) over the full rollout, population std (unbiased=False)."""
    return (adv - adv.mean()) / (adv.std(unbiased=False) + eps)


def ppo_loss(logits, actions, logp_old, adv, value, ret, l_sem, spec: SpecV0):
    """a9.  Returns dict of scalar tensors; `total` carries the autograd graph."""
    logp_all = torch.log_softmax(logits.float(), dim=-1)
    logp = logp_all.gather(-1, actions.view(-1, 1)).squeeze(-1)
    ratio = torch.exp(logp - logp_old)
    s1 = ratio * adv
    s2 = torch.clamp(ratio, 1.0 - spec.clip_eps, 1.0 + spec.clip_eps) * adv
    l_pi = -torch.min(s1, s2).mean()
    l_v = 0.5 * (value.float() - ret).pow(2).mean()
    ent = -(logp_all.exp() * logp_all).sum(-1).mean()
    total = l_pi + spec.c_v * l_v - spec.c_ent * ent + spec.c_sem * l_sem
    with torch.no_grad():
        kl = (logp_old - logp).mean()
        clipfrac = ((ratio - 1.0).abs() > spec.clip_eps).float().mean()
    return {"total": total, "pi": l_pi, "v": l_v, "ent": ent, "l_sem": l_sem,
            "kl": kl, "clipfrac": clipfrac}


def minibatch_permutation(n: int, generator: torch.Generator) -> torch.Tensor:
    """a10: int64 permutation from a seeded HOST generator (bit-exact by construction)."""
    return torch.randperm(n, generator=generator)


class PPOAgent:
    """update(rollout) -> stats.  rollout: dict of tensors
        cam[T,N,3,H,W] bev[T,N,1,h,w] actions[T,N] logp[T,N] values[T,N]
        rewards[T,N] dones[T,N] last_value[N]"""

    def __init__(self, policy, spec: SpecV0 = SpecV0(), *, epochs: int = 1,
                 minibatch: int = 4096, snr_db: float = 10.0, channel: str = "awgn",
                 seed: int = 1234):
        self.policy = policy
        self.spec = spec
        self.epochs = epochs
        self.minibatch = minibatch
        self.snr_db = snr_db
        self.channel = channel
        self.seed = seed
        self.step_count = 0
        self.gen = torch.Generator().manual_seed(seed)
        self.opt = torch.optim.Adam(policy.parameters(), lr=spec.lr,
                                    betas=(spec.beta1, spec.beta2), eps=spec.adam_eps)
        self.last_perms = []          # permutations / per-minibatch statistics of the last update() (parity tests)
        self.last_step_stats = []

    def minibatch_step(self, cam, bev, actions, logp_old, adv, values_unused, ret,
                       sample_offset: int = 0):
        """One optimizer step on one minibatch (the unit bench.py counts)."""
        sp = self.spec
        logits, value, aux = self.policy(cam, bev, self.snr_db, channel=self.channel,
                                         seed=self.seed + self.step_count,
                                         sample_offset=sample_offset)
        out = ppo_loss(logits, actions, logp_old, adv, value, ret, aux["l_sem"], sp)
        self.opt.zero_grad(set_to_none=True)
        out["total"].backward()
        self.opt.step()
        self.step_count += 1
        return {k: float(v.detach()) for k, v in out.items()}

    def update(self, rollout: dict) -> dict:
        sp = self.spec
        T, N = rollout["rewards"].shape
        adv, ret = gae(rollout["rewards"], rollout["values"], rollout["dones"],
                       rollout["last_value"], sp.gamma, sp.lam)
        adv = normalize_adv(adv, sp.adv_eps)
        n = T * N
        flat = lambda x: x.reshape(n, *x.shape[2:])
        cam, bev = flat(rollout["cam"]), flat(rollout["bev"])
        act, lp, val = flat(rollout["actions"]), flat(rollout["logp"]), flat(rollout["values"])
        adv, ret = adv.reshape(n), ret.reshape(n)
        mb = min(self.minibatch, n)
        stats, cnt = {}, 0
        self.last_perms, self.last_step_stats = [], []
        for _ in range(self.epochs):
            perm = minibatch_permutation(n, self.gen)
            self.last_perms.append(perm)
            for i in range(0, n - mb + 1, mb):
                idx = perm[i:i + mb].to(cam.device)
                s = self.minibatch_step(cam[idx], bev[idx], act[idx], lp[idx], adv[idx],
                                        val[idx], ret[idx])
                self.last_step_stats.append(s)
                for k, v in s.items():
                    stats[k] = stats.get(k, 0.0) + v
                cnt += 1
        return {k: v / max(cnt, 1) for k, v in stats.items()}
