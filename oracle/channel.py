"""Power-normalised noisy channel — eager PyTorch restatement of SPEC v0 rows a3/a4.

Reference citation: no upstream code exists (/root/reference/README.md:1-2 only);
this follows SURVEY.md §8 rows a3 (power_normalize) and a4 (awgn, rayleigh).
All functions are differentiable torch code so autograd provides the oracle gradients.
"""
from __future__ import annotations

import numpy as np
import torch

from . import philox


def power_normalize(z: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """z[B, 2, k] -> unit mean power per (sample, modality): z * rsqrt(mean(z^2) + eps)."""
    p = z.pow(2).mean(dim=-1, keepdim=True)
    return z * torch.rsqrt(p + eps)


def quantize_ste(zn: torch.Tensor, bits: int, clip: float) -> torch.Tensor:
    """b-bit quantised symbols (no upstream code, /root/reference/README.md:1-2): mid-rise uniform
    quantiser with L = 2^bits levels on [-clip, clip], q = (clamp(floor(zn / delta), -L/2, L/2 - 1) + 0.5) delta,
    delta = 2 clip / L.  Gradient: clipped straight-through estimator (1 inside [-clip, clip], 0 outside)."""
    L = 2 ** int(bits)
    delta = 2.0 * clip / L
    q = (torch.clamp(torch.floor(zn * (L / (2.0 * clip))), -L / 2, L / 2 - 1) + 0.5) * delta
    inside = (zn.abs() <= clip).to(zn.dtype)
    passed = zn * inside
    return passed + (q - passed).detach()


def snr_to_sigma(snr_db, batch: int, device, dtype) -> torch.Tensor:
    """sigma[B,1,1] with sigma^2 = 10^(-snr_db/10) per real dimension."""
    s = torch.as_tensor(snr_db, dtype=torch.float32, device=device)
    if s.dim() == 0:
        s = s.expand(batch)
    sigma = torch.pow(torch.tensor(10.0, dtype=torch.float32, device=device), -s / 20.0)
    return sigma.to(dtype).view(batch, 1, 1)


def make_noise(seed: int, batch: int, k: int, sample_offset: int = 0) -> torch.Tensor:
    return torch.from_numpy(philox.noise_normal(seed, batch, k, sample_offset))


def make_fading(seed: int, batch: int, sample_offset: int = 0) -> torch.Tensor:
    return torch.from_numpy(philox.fading_coeff(seed, batch, sample_offset))


def awgn(zn: torch.Tensor, sigma: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """y = zn + sigma * n."""
    return zn + sigma * noise.to(zn.dtype)


def rayleigh(zn: torch.Tensor, sigma: torch.Tensor, noise: torch.Tensor,
             h: torch.Tensor) -> torch.Tensor:
    """Block Rayleigh fading with MMSE equalisation.
    reals are paired (even, odd) -> complex s; r = h s + sigma n; y = conj(h) r / (|h|^2 + sigma^2).
    h: [B, 2, 2] = (re, im) per (sample, modality)."""
    B, M, k = zn.shape
    s = zn.reshape(B, M, k // 2, 2)
    n = noise.to(zn.dtype).reshape(B, M, k // 2, 2)
    hr = h[..., 0].to(zn.dtype).view(B, M, 1)
    hi = h[..., 1].to(zn.dtype).view(B, M, 1)
    sr, si = s[..., 0], s[..., 1]
    sg = sigma.view(B, 1, 1)
    rr = hr * sr - hi * si + sg * n[..., 0]
    ri = hr * si + hi * sr + sg * n[..., 1]
    den = hr * hr + hi * hi + sg * sg
    yr = (hr * rr + hi * ri) / den
    yi = (hr * ri - hi * rr) / den
    return torch.stack([yr, yi], dim=-1).reshape(B, M, k)


def rician_fading(h0: torch.Tensor, k_factor: float) -> torch.Tensor:
    """CN(0,1) block-fading draws h0[B,2,2] -> Rician coefficients with linear K-factor `k_factor`:
    h = sqrt(K/(K+1)) + sqrt(1/(K+1)) h0 (line of sight on the real axis, E|h|^2 = 1).  K = 0 returns h0 (Rayleigh).
    SURVEY.md §8f rank 3; no upstream code exists (/root/reference/README.md:1-2)."""
    k = torch.tensor(float(k_factor), dtype=torch.float32)
    los, sc = torch.sqrt(k / (k + 1.0)), torch.sqrt(1.0 / (k + 1.0))
    h = h0.to(torch.float32) * sc
    h = torch.stack([h[..., 0] + los, h[..., 1]], dim=-1)
    return h.to(h0.dtype)


def channel_forward(z: torch.Tensor, snr_db, kind: str, *, seed=None, noise=None, fading=None,
                    sample_offset: int = 0, eps: float = 1e-12, rician_k: float = 4.0, quant_bits: int = 0,
                    quant_clip: float = 4.0, sample_idx=None):
    """z[B,2,k] -> (y[B,2,k], zn).  kind in {"awgn", "rayleigh", "rician", "none"}.  An explicit `fading` tensor is the
    channel coefficient h itself for both fading kinds.  quant_bits > 0: the channel input is quantize_ste(zn).
    sample_idx (int64 [B], optional): global sample index of every row for the noise counters (ragged data-parallel
    slices); default sample_offset + b."""
    B, M, k = z.shape
    zn = power_normalize(z, eps)
    tx = quantize_ste(zn, quant_bits, quant_clip) if quant_bits > 0 else zn
    if kind == "none":
        return tx, zn
    if sample_idx is not None and noise is None:
        ids = [int(i) for i in sample_idx]
        noise = torch.cat([make_noise(seed, 1, k, i) for i in ids]).to(z.device)
        if kind in ("rayleigh", "rician") and fading is None:
            fading = torch.cat([make_fading(seed, 1, i) for i in ids])
            if kind == "rician":
                fading = rician_fading(fading, rician_k)
            fading = fading.to(z.device)
    sigma = snr_to_sigma(snr_db, B, z.device, z.dtype)
    if noise is None:
        if seed is None:
            raise ValueError("channel needs seed= or noise=")
        noise = make_noise(seed, B, k, sample_offset).to(z.device)
    if kind == "awgn":
        return awgn(tx, sigma, noise), zn
    if kind == "rayleigh":
        if fading is None:
            if seed is None:
                raise ValueError("rayleigh channel needs seed= or fading=")
            fading = make_fading(seed, B, sample_offset).to(z.device)
        return rayleigh(tx, sigma, noise, fading), zn
    if kind == "rician":                                      # same receiver, line-of-sight component added to h
        if fading is None:
            if seed is None:
                raise ValueError("rician channel needs seed= or fading=")
            fading = rician_fading(make_fading(seed, B, sample_offset), rician_k).to(z.device)
        return rayleigh(tx, sigma, noise, fading), zn
    raise ValueError(f"unknown channel kind {kind!r}")


def empirical_snr_db(zn: torch.Tensor, y: torch.Tensor) -> float:
    """10 log10(P_signal / P_noise) for an AWGN output (test helper)."""
    n = (y - zn).double()
    return float(10.0 * np.log10(zn.double().pow(2).mean().item() / n.pow(2).mean().item()))
