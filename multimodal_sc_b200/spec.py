"""SPEC v0 — the frozen definition of the hot path ("SPEC v0 constants").

The upstream tree (/root/reference/README.md:1-2) holds a title and no code, so the
shapes and hyper-parameters of the path are fixed HERE, once, and shared by the
oracle (`oracle/`), the native sm_100a path (`multimodal_sc_b200/`) and `bench.py`.

Revisions relative to SURVEY.md §8 (allowed "once in PR1, then frozen"):
  * u32 -> uniform conversion is ((x >> 9) + 0.5) * 2**-23 instead of
    (x + 0.5) * 2**-32, so that it is exactly representable in fp32 and therefore
    bit-identical between numpy, torch-CPU and the CUDA kernels.
  * power normalisation is z * rsqrt(mean(z^2) + 1e-12) (eps inside the sqrt, on
    the mean rather than the sum).
  * GELU is the tanh form 0.5 x (1 + tanh(0.79788456 (x + 0.044715 x^3))) (one MUFU.TANH per
    element in the GEMM epilogue instead of an erf polynomial; the D=256 GEMMs leave ~10
    issue slots per output element before the epilogue, not the tensor pipe, bounds them).
  * every encoder and the decoder end with a LayerNorm (pre-LN stacks need one);
    the semantic target of L_sem is that final-LN encoder feature.

Round-2 OPTIONS (every default reproduces the frozen SPEC v0 bit for bit):
  * gelu = "erf": the exact GELU x Phi(x) that SURVEY.md §8 originally froze, for interchange with an upstream
    `nn.GELU()` model should one ever be published (direct-store GEMM epilogue; measured cost in profiles/);
  * snr_cond: SNR-conditioned encoders, a per-sample embedding W tanh(s u + v), s = snr_db / 20, added to every
    token next to the positional embedding (parameters {cam,bev}_enc.snr.{u,v,w});
  * quant_bits > 0: the normalised symbols pass a 2^bits-level uniform quantiser on [-quant_clip, quant_clip]
    before the channel, clipped straight-through gradient.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict


@dataclass(frozen=True)
class SpecV0:
    # model
    d_model: int = 256
    n_heads: int = 4
    mlp_ratio: int = 4
    enc_layers: int = 2          # per modality
    dec_layers: int = 2
    cam_ch: int = 3
    cam_hw: int = 128
    cam_patch: int = 16          # -> 64 tokens, K = 768
    bev_ch: int = 1
    bev_hw: int = 64
    bev_patch: int = 8           # -> 64 tokens, K = 64
    c_sym: int = 16              # reals per token -> k = 1024 reals / modality / sample
    n_actions: int = 9
    ln_eps: float = 1e-5
    init_std: float = 0.02
    # channel
    pnorm_eps: float = 1e-12
    rician_k: float = 4.0          # linear K-factor (6 dB) of the "rician" channel kind
    # round-2 options (defaults = frozen SPEC v0)
    gelu: str = "tanh"             # "tanh" | "erf"
    snr_cond: bool = False         # SNR-conditioned encoders
    snr_embed_dim: int = 64
    quant_bits: int = 0            # 0 = continuous symbols
    quant_clip: float = 4.0        # a power of two keeps the quantiser step exact in fp32
    # PPO
    gamma: float = 0.99
    lam: float = 0.95
    clip_eps: float = 0.2
    c_v: float = 0.5
    c_ent: float = 0.01
    c_sem: float = 1.0
    lr: float = 3e-4
    beta1: float = 0.9
    beta2: float = 0.999
    adam_eps: float = 1e-8
    adv_eps: float = 1e-8
    rollout_t: int = 128

    # ---- derived -------------------------------------------------------
    @property
    def d_head(self) -> int:
        return self.d_model // self.n_heads

    @property
    def cam_tokens(self) -> int:
        return (self.cam_hw // self.cam_patch) ** 2

    @property
    def bev_tokens(self) -> int:
        return (self.bev_hw // self.bev_patch) ** 2

    @property
    def cam_k(self) -> int:
        return self.cam_ch * self.cam_patch ** 2

    @property
    def bev_k(self) -> int:
        return self.bev_ch * self.bev_patch ** 2

    @property
    def dec_tokens(self) -> int:
        return self.cam_tokens + self.bev_tokens

    @property
    def k_sym(self) -> int:
        """real channel uses per modality per sample"""
        assert self.cam_tokens == self.bev_tokens
        return self.cam_tokens * self.c_sym

    @property
    def d_mlp(self) -> int:
        return self.d_model * self.mlp_ratio

    def to_dict(self) -> dict:
        return asdict(self)


# Philox4x32-10 constants (Salmon et al., SC'11) — shared so oracle and tests agree.
PHILOX_M0 = 0xD2511F53
PHILOX_M1 = 0xCD9E8D57
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
STREAM_NOISE = 0
STREAM_FADING = 1

# fwd MFLOP per sample by piece, used by bench.py for the tensor roofline
def fwd_flops_per_sample(spec: SpecV0 = SpecV0()) -> float:
    D, S, Sd, R = spec.d_model, spec.cam_tokens, spec.dec_tokens, spec.mlp_ratio

    def block(s):
        lin = 2 * s * D * (3 * D + D + 2 * R * D)
        att = 2 * 2 * s * s * D
        return lin + att

    f = 2 * S * spec.cam_k * D + 2 * spec.bev_tokens * spec.bev_k * D
    f += 2 * spec.enc_layers * block(S)
    f += 2 * 2 * S * D * spec.c_sym            # symbol projections
    f += 2 * Sd * spec.c_sym * D               # decoder in-proj
    f += spec.dec_layers * block(Sd)
    f += 2 * Sd * D * D                        # recon head
    f += 2 * D * D + 2 * D * (spec.n_actions + 1)
    return float(f)
