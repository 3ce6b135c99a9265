"""Eager-PyTorch fp32 restatement of the SPEC v0 model (rows a1, a2, a5, a6, a7).

Reference citation: no upstream model exists (/root/reference/README.md:1-2 only —
"Semantic-Aware Multi-Modal Reinforcement Learning for Autonomous Driving"); the layer
stack follows SURVEY.md §3.2(A) and §8 rows a1/a2/a5/a6/a7.  This is the "PR1 ref model"
of BASELINE.json configs[0].  state_dict keys/shapes are shared with the native
`multimodal_sc_b200.model.SemComPolicy` so weights load both ways.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from multimodal_sc_b200.spec import SpecV0
from . import channel as ch


def patchify(img: torch.Tensor, p: int) -> torch.Tensor:
    """[B,C,H,W] -> [B, (H/p)*(W/p), C*p*p]; token order (ty, tx), feature order (c, py, px)."""
    B, C, H, W = img.shape
    x = img.reshape(B, C, H // p, p, W // p, p)
    x = x.permute(0, 2, 4, 1, 3, 5)
    return x.reshape(B, (H // p) * (W // p), C * p * p)


class Block(nn.Module):
    """pre-LN transformer block: x += proj(attn(ln1 x)); x += fc2(gelu(fc1(ln2 x))); GELU in the tanh form (SPEC v0) or
    exact (spec.gelu == "erf")."""

    def __init__(self, spec: SpecV0):
        super().__init__()
        D = spec.d_model
        self.h = spec.n_heads
        self.gelu_approx = "tanh" if spec.gelu == "tanh" else "none"
        self.ln1 = nn.LayerNorm(D, eps=spec.ln_eps)
        self.qkv = nn.Linear(D, 3 * D)
        self.proj = nn.Linear(D, D)
        self.ln2 = nn.LayerNorm(D, eps=spec.ln_eps)
        self.fc1 = nn.Linear(D, spec.d_mlp)
        self.fc2 = nn.Linear(spec.d_mlp, D)

    def attention(self, hx: torch.Tensor) -> torch.Tensor:
        B, S, D = hx.shape
        dh = D // self.h
        qkv = self.qkv(hx).view(B, S, 3, self.h, dh)
        q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))     # [B,H,S,dh]
        att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(dh), dim=-1)
        return (att @ v).transpose(1, 2).reshape(B, S, D)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x + self.proj(self.attention(self.ln1(x)))
        x = x + self.fc2(F.gelu(self.fc1(self.ln2(x)), approximate=self.gelu_approx))
        return x


class SnrEmbed(nn.Module):
    """SNR conditioning (no upstream code): emb[b] = w tanh(s_b u + v), s_b = snr_db[b] / 20."""

    def __init__(self, spec: SpecV0):
        super().__init__()
        E, D = spec.snr_embed_dim, spec.d_model
        self.u = nn.Parameter(torch.zeros(E))
        self.v = nn.Parameter(torch.zeros(E))
        self.w = nn.Parameter(torch.zeros(D, E))

    def forward(self, snr_db, batch: int, device) -> torch.Tensor:
        s = torch.as_tensor(snr_db, dtype=torch.float32, device=device)
        if s.dim() == 0:
            s = s.expand(batch)
        s = s * 0.05
        return torch.tanh(s[:, None] * self.u + self.v) @ self.w.t()          # [B, D]


class Encoder(nn.Module):
    """patch-embed + pos -> L blocks -> LayerNorm (feature t) -> Linear(D, c_sym) (symbols z)."""

    def __init__(self, spec: SpecV0, ch_in: int, hw: int, patch: int):
        super().__init__()
        self.patch = patch
        tokens = (hw // patch) ** 2
        self.patch_embed = nn.Linear(ch_in * patch * patch, spec.d_model)
        self.pos = nn.Parameter(torch.zeros(tokens, spec.d_model))
        self.blocks = nn.ModuleList([Block(spec) for _ in range(spec.enc_layers)])
        self.norm = nn.LayerNorm(spec.d_model, eps=spec.ln_eps)
        self.sym = nn.Linear(spec.d_model, spec.c_sym)
        self.snr = SnrEmbed(spec) if spec.snr_cond else None

    def forward(self, img: torch.Tensor, snr_db=None):
        x = self.patch_embed(patchify(img, self.patch)) + self.pos
        if self.snr is not None:
            x = x + self.snr(snr_db, img.shape[0], img.device)[:, None, :]
        for b in self.blocks:
            x = b(x)
        t = self.norm(x)                                   # [B,S,D]  semantic feature
        z = self.sym(t)                                    # [B,S,c_sym]
        return z.reshape(z.shape[0], -1), t


class CameraEncoder(Encoder):
    def __init__(self, spec: SpecV0):
        super().__init__(spec, spec.cam_ch, spec.cam_hw, spec.cam_patch)


class LidarEncoder(Encoder):
    def __init__(self, spec: SpecV0):
        super().__init__(spec, spec.bev_ch, spec.bev_hw, spec.bev_patch)


class FusionDecoder(nn.Module):
    """received symbols of both modalities -> joint self-attention -> pooled feature + recon."""

    def __init__(self, spec: SpecV0):
        super().__init__()
        D = spec.d_model
        self.c_sym = spec.c_sym
        self.in_proj = nn.Linear(spec.c_sym, D)
        self.pos = nn.Parameter(torch.zeros(spec.dec_tokens, D))     # position + modality embedding
        self.blocks = nn.ModuleList([Block(spec) for _ in range(spec.dec_layers)])
        self.norm = nn.LayerNorm(D, eps=spec.ln_eps)
        self.recon = nn.Linear(D, D)

    def forward(self, y: torch.Tensor):
        B = y.shape[0]
        x = self.in_proj(y.reshape(B, -1, self.c_sym)) + self.pos    # [B,128,D] (cam tokens first)
        for b in self.blocks:
            x = b(x)
        x = self.norm(x)
        return x.mean(dim=1), self.recon(x)


class ActorCritic(nn.Module):
    def __init__(self, spec: SpecV0):
        super().__init__()
        D = spec.d_model
        self.trunk = nn.Linear(D, D)
        self.head = nn.Linear(D, spec.n_actions + 1)       # rows 0..A-1 = logits, row A = value

    def forward(self, f: torch.Tensor):
        o = self.head(torch.tanh(self.trunk(f)))
        return o[:, :-1], o[:, -1]


def semantic_distortion(recon: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """a6: mean((recon - stop_grad(target))^2)."""
    return (recon - target.detach()).pow(2).mean()


def sample_action(logits: torch.Tensor, u: torch.Tensor):
    """a7: inverse-CDF sampling from a SUPPLIED uniform u[B].

    action = #{j : cdf_j <= u}, clamped to A-1, with the cdf accumulated in index order in
    float64 from the fp32 logits (float64 so that CPU and GPU agree on the integer result
    unless u lies within ~1e-15 of a boundary).  Returns (action int64, logp fp32)."""
    l64 = logits.double()
    p = torch.softmax(l64, dim=-1)
    cdf = torch.cumsum(p, dim=-1)
    a = (cdf <= u.double().unsqueeze(-1)).sum(dim=-1).clamp_(max=logits.shape[-1] - 1)
    logp = torch.log_softmax(logits.float(), dim=-1).gather(-1, a.unsqueeze(-1)).squeeze(-1)
    return a.to(torch.int64), logp


class SemComPolicy(nn.Module):
    """enc(cam), enc(bev) -> power-norm + channel -> fusion decoder -> actor-critic."""

    def __init__(self, spec: SpecV0 = SpecV0()):
        super().__init__()
        self.spec = spec
        self.cam_enc = CameraEncoder(spec)
        self.bev_enc = LidarEncoder(spec)
        self.dec = FusionDecoder(spec)
        self.ac = ActorCritic(spec)
        self.reset_parameters()

    def reset_parameters(self, std: float | None = None, seed: int | None = 0):
        std = self.spec.init_std if std is None else std
        g = torch.Generator().manual_seed(seed) if seed is not None else None
        with torch.no_grad():
            for name, p in self.named_parameters():
                if name.endswith("bias"):
                    p.zero_()
                elif ".ln" in name or "norm" in name:
                    p.fill_(1.0)                            # LayerNorm weight
                else:
                    w = torch.empty(p.shape).normal_(0.0, 1.0, generator=g).clamp_(-2.0, 2.0)
                    p.copy_(w * std)

    def forward(self, cam, bev, snr_db, *, channel: str = "awgn", seed=None, noise=None,
                fading=None, sample_offset: int = 0, rician_k=None, sample_idx=None):
        z_c, t_c = self.cam_enc(cam, snr_db)
        z_b, t_b = self.bev_enc(bev, snr_db)
        z = torch.stack([z_c, z_b], dim=1)                  # [B,2,k]
        y, zn = ch.channel_forward(z, snr_db, channel, seed=seed, noise=noise, fading=fading,
                                   sample_offset=sample_offset, eps=self.spec.pnorm_eps,
                                   rician_k=self.spec.rician_k if rician_k is None else rician_k,
                                   quant_bits=self.spec.quant_bits, quant_clip=self.spec.quant_clip,
                                   sample_idx=sample_idx)
        f, recon = self.dec(y)
        logits, value = self.ac(f)
        l_sem = semantic_distortion(recon, torch.cat([t_c, t_b], dim=1))
        return logits, value, {"l_sem": l_sem, "z": zn, "y": y}

    @torch.no_grad()
    def act(self, cam, bev, snr_db, u, **kw):
        logits, value, _ = self.forward(cam, bev, snr_db, **kw)
        a, logp = sample_action(logits, u)
        return a, logp, value
