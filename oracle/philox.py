"""Philox4x32-10 counter RNG + the SPEC v0 noise layout (numpy, exact integer arithmetic).

Reference citation: none exists upstream (/root/reference/README.md:1-2 only); the
algorithm is Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3"
(SC'11), Philox4x32 with 10 rounds, pinned by its published known-answer vectors
(SURVEY.md §4.2).  SPEC v0 row a4 fixes how counters map to channel symbols:

    flat real-symbol index i in [B_global, 2, k]   (k = 1024)
    counter = (lo32(i >> 2), stream, hi32(i >> 2), 0), key = (seed_lo, seed_hi)
    lane    = i & 3  ->  word `lane` of the 4-word output
    u       = ((x >> 9) + 0.5) * 2**-23            (exact in fp32, strictly inside (0,1))
    Box-Muller pairs lanes (0,1) and (2,3): n_even = r cos(2 pi u_odd), n_odd = r sin(2 pi u_odd),
    r = sqrt(-2 ln u_even)
    stream 0 = additive noise, stream 1 = block-fading coefficient h (one call per
    (sample, modality): counter = (j, 1, 0, 0), j = b*2 + m; lanes 0,1 -> (h_re, h_im)/sqrt(2))
"""
from __future__ import annotations

import numpy as np

from multimodal_sc_b200.spec import (PHILOX_M0, PHILOX_M1, PHILOX_W0, PHILOX_W1,
                                     STREAM_FADING, STREAM_NOISE)

_U32 = np.uint32
_U64 = np.uint64
_MASK = _U64(0xFFFFFFFF)


def _mulhilo(a: int, b: np.ndarray):
    p = _U64(a) * b.astype(_U64)
    return (p >> _U64(32)).astype(_U32), (p & _MASK).astype(_U32)


def philox4x32_10(ctr: np.ndarray, key) -> np.ndarray:
    """ctr: uint32[..., 4]; key: (k0, k1) ints. Returns uint32[..., 4]."""
    ctr = np.asarray(ctr, dtype=_U32)
    c0, c1, c2, c3 = (ctr[..., i].copy() for i in range(4))
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for r in range(10):
        if r > 0:
            k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
            k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
        hi0, lo0 = _mulhilo(PHILOX_M0, c0)
        hi1, lo1 = _mulhilo(PHILOX_M1, c2)
        c0, c1, c2, c3 = hi1 ^ c1 ^ _U32(k0), lo1, hi0 ^ c3 ^ _U32(k1), lo0
    return np.stack([c0, c1, c2, c3], axis=-1)


def seed_to_key(seed: int):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def u32_to_uniform(x: np.ndarray) -> np.ndarray:
    """((x >> 9) + 0.5) * 2**-23 as float64 (value is exactly representable in fp32)."""
    return ((x >> _U32(9)).astype(np.float64) + 0.5) * (2.0 ** -23)


def box_muller(u_a: np.ndarray, u_b: np.ndarray):
    """float64 Box-Muller; callers cast to fp32 at the end."""
    r = np.sqrt(-2.0 * np.log(u_a))
    th = 2.0 * np.pi * u_b
    return r * np.cos(th), r * np.sin(th)


def philox_words(seed: int, n_quads: int, stream: int, quad_offset: int = 0) -> np.ndarray:
    """uint32[n_quads, 4]: output words for counters quad_offset .. quad_offset+n_quads-1."""
    q = np.arange(quad_offset, quad_offset + n_quads, dtype=np.uint64)
    ctr = np.zeros((n_quads, 4), dtype=_U32)
    ctr[:, 0] = (q & _MASK).astype(_U32)
    ctr[:, 1] = _U32(stream)
    ctr[:, 2] = (q >> _U64(32)).astype(_U32)
    return philox4x32_10(ctr, seed_to_key(seed))


def noise_normal(seed: int, batch: int, k: int, sample_offset: int = 0) -> np.ndarray:
    """Standard-normal noise float32[batch, 2, k] for global samples
    sample_offset .. sample_offset+batch-1 (SPEC v0 a4, stream 0)."""
    per = 2 * k
    assert per % 4 == 0
    n_quads = batch * per // 4
    w = philox_words(seed, n_quads, STREAM_NOISE, quad_offset=sample_offset * per // 4)
    u = u32_to_uniform(w)
    n0, n1 = box_muller(u[:, 0], u[:, 1])
    n2, n3 = box_muller(u[:, 2], u[:, 3])
    out = np.stack([n0, n1, n2, n3], axis=-1).reshape(batch, 2, k)
    return out.astype(np.float32)


def fading_coeff(seed: int, batch: int, sample_offset: int = 0) -> np.ndarray:
    """h ~ CN(0,1) as float32[batch, 2, 2] = (re, im) per (sample, modality) (stream 1)."""
    n = batch * 2
    w = philox_words(seed, n, STREAM_FADING, quad_offset=sample_offset * 2)
    u = u32_to_uniform(w)
    hr, hi = box_muller(u[:, 0], u[:, 1])
    h = np.stack([hr, hi], axis=-1) / np.sqrt(2.0)
    return h.reshape(batch, 2, 2).astype(np.float32)
