// gemm_tcgen05.cu — persistent, warp-specialised bf16 GEMM for sm_100a.
//
//   acc[m,n] = sum_k A(m,k) B(n,k), fp32 accumulation in TMEM, fused epilogue (see
//   include/semcom_sm100.h: sc_gemm_bf16).  Implements every dense contraction of SPEC v0
//   rows a1/a2/a5 and their dgrad / wgrad forms.
//
// Structure (one CTA per SM, 384 threads):
//   warp 0 / lane 0 : TMA producer     — cp.async.bulk.tensor.2d into a 3-stage smem ring
//   warp 1 / lane 0 : MMA issuer       — tcgen05.mma.cta_group::1.kind::f16, 128x256x16
//   warp 2          : TMEM allocator   — 512 columns = two 128x256 fp32 accumulators
//   warps 4..11     : epilogue (two warpgroups, one per accumulator stage)         — tcgen05.ld 32x32b, fused math; bf16 tiles staged in a per-warp
//                                        SWIZZLE_128B smem ring and moved by TMA store / TMA load
//                                        (fp32 / ragged outputs: direct 16-byte stores)
// Three mbarrier pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue).
// The accumulator is double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Operand layouts in shared memory (all SWIZZLE_128B, written by TMA):
//   K-major  operand (stored [rows, K]): tile = rows x 64 k  -> rows x 128 B, 8-row groups of
//            1024 B (SBO = 1024); successive UMMA_K=16 slices advance the start by 32 B.
//   MN-major operand (stored [K, cols]): tile = 64-column chunks, each 64 k-rows x 128 B
//            (8192 B, LBO = 8192 between chunks, SBO = 1024 between 8-k-row groups);
//            successive UMMA_K slices advance the start by 16 rows = 2048 B.
#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;                 // 16384
// CTA2 = 0: one CTA per tile (cta_group::1), B tile 256 x 64 per stage (48 KB stages).
// CTA2 = 2: a cluster of two INDEPENDENT CTAs (cta_group::1, own accumulators and epilogues) that work on the same N tile
//           of two consecutive 128-row blocks and share the B (weight) tile of every stage: each CTA loads half of it
//           and multicasts it into both shared memories; a stage is free when both MMA warps have committed.  Halves
//           the weight traffic out of L2 (2/3 of the operand bytes of a K = 256 tile) without the accumulator hand-off
//           that couples the epilogues of a cta_group::2 pair.
// CTA2 = 1: a CTA pair (cluster of 2, cta_group::2) computes a 256 x 256 tile; each CTA stages its own 128 A rows
//           and HALF of the B tile (128 x 64) -> 32 KB stages and half the B shared-memory traffic per SM.
// NWG   = epilogue warpgroups (1 or 2): two warps per scheduler hide each other's dependency stalls.
// SLOTS = 4 KB staging slots per epilogue warp (2 or 4).  4 slots hold a whole 32 x 256 tile row block, so the
//         bf16 input tile of the NEXT tile is prefetched a full tile ahead and two-output epilogues double-buffer.
// The pipeline gets every stage that still fits next to the staging ring and the bias tiles.
// wait accounting (gemm_dbg bit 512) is compiled in only with -DSC_GEMM_ACCT (SC_GEMM_ACCT=1 python -m
// multimodal_sc_b200.build --force): its counters cost registers the 384-thread variants do not have to spare
#ifdef SC_GEMM_ACCT
constexpr bool kAcct = true;
#else
constexpr bool kAcct = false;
#endif
constexpr int SMEM_LIMIT = 232448;                   // 227 KB opt-in maximum per CTA
constexpr int STATIC_SMEM_RESERVE = 1024;            // mbarriers etc. (static __shared__), rounded up
constexpr int EPI_SLOT_BYTES = 32 * 128;             // 32 rows x 64 bf16 columns, SWIZZLE_128B
constexpr int BIAS_TILE_BYTES = BN * 4;              // fp32 bias of the current N tile, one copy per epilogue warpgroup
template <int CTA2, int NWG, int SLOTS> struct Geo {
  static constexpr int B_ROWS = (CTA2 == 1) ? BN / 2 : BN;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RING = NWG * 4 * SLOTS * EPI_SLOT_BYTES;
  // one bias tile per warpgroup; the 1-warpgroup / 4-slot geometry also parks LayerNorm gamma and beta (fused LN)
  // (the 2-warpgroup / 2-slot geometry runs the fused LayerNorm with the row split between the warpgroups: columns
  //  0-127 / 128-255, partial row statistics exchanged through LN_XCHG bytes of shared memory, double-buffered by tile)
  static constexpr int LN_TILES = (CTA2 != 1 && ((NWG == 1 && SLOTS == 4) || (NWG == 2 && SLOTS == 2))) ? 2 : 0;
  static constexpr int LN_XCHG = (CTA2 != 1 && NWG == 2 && SLOTS == 2) ? 2 * 2 * BM * 8 : 0;
  static constexpr int BIAS = (NWG + LN_TILES) * BIAS_TILE_BYTES + LN_XCHG;
  static constexpr int STAGES_FIT = (SMEM_LIMIT - STATIC_SMEM_RESERVE - RING - BIAS) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 6 ? 6 : STAGES_FIT;
  static constexpr int SMEM = STAGES * STAGE_BYTES;
  static constexpr int THREADS = 128 + NWG * 128;
  static constexpr int TOTAL = SMEM + RING + BIAS;    // dynamic shared memory (base is 1 KB aligned)
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
};
constexpr int CHUNK_BYTES = 64 * BK * 2;             // 8192: one 64-wide MN-major chunk
constexpr int MAX_EPI_WARPS = 8;
constexpr int MAX_SLOTS = 4;
// staged-epilogue modes
constexpr int EPI_GENERIC = 0;                       // direct register -> global stores (fp32 / ragged outputs)
constexpr int EPI_OUT1 = 1;                          // bf16 out via smem + TMA store
constexpr int EPI_OUT2 = 2;                          // bf16 out + bf16 pre-activation copy (two TMA stores per chunk)
constexpr int EPI_IN = 3;                            // bf16 out, one bf16 input tile (residual or GELU' operand) by TMA load
// compile-time specialisations of the staged epilogue math (small hot loops -> fit the L0 I-cache)
constexpr int KIND_GEN = 0;                          // everything decided at run time
constexpr int KIND_PLAIN = 1;                        // bf16(acc)
constexpr int KIND_BIAS = 2;                         // + bias
constexpr int KIND_FC1 = 3;                          // + bias, pre-activation copy, tanh-GELU
constexpr int KIND_RESID = 4;                        // + bias + residual tile
constexpr int KIND_DGELU = 5;                        // * GELU'(pre-activation tile)
constexpr int KIND_RESID_LN = 6;                     // + bias + residual tile, then LayerNorm of the row as a second output
constexpr uint32_t TMEM_COLS = 512;

struct GemmParams {
  int64_t M, N, K;
  int tiles_m, tiles_n, ksplit, kb_total, kb_per_split;
  // epilogue
  const float* bias;
  const float* rowadd;
  uint32_t rowadd_period;
  const float* rowgrp;
  uint32_t rowgrp_size;
  float rowgrp_scale;
  float alpha;
  const bf16* resid;
  const bf16* aux_in;
  bf16* aux_out;
  bf16* out_bf16;
  float* out_f32;
  int64_t ldo;
  int act;
  int aux_in_erf;
  int atomic;
  int vec_ok;
  int epi_mode;
  int epi_kind;
  float* colsum;                // a_mn = 1 only: colsum[m] += sum_k A(m,k) via an extra MMA against an all-ones tile
  int in_row_period;            // > 0: the TMA input tile is a [period, N] table indexed by row % period
  // fused LayerNorm of the output row (N == 256 only): second bf16 output through tmAux, statistics to ln_mean/ln_rstd
  const float* ln_gamma;
  const float* ln_beta;
  float* ln_mean;
  float* ln_rstd;
  float ln_eps;
  // UMMA descriptor fields (bytes); defaults set by the host, overridable for bring-up
  uint32_t a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep;
  // skinny problems (one N tile with N < 256, one M tile with M <= 64): only the live part of the operand tiles is
  // loaded and the MMA is issued with the narrow N — TMA boxes that are mostly out of bounds are slow to zero-fill
  int n_mma;                    // MMA N = rows of the K-major B box (multiple of 16, <= 256)
  int n_chunks;                 // 64-column chunks of an MN-major B tile that are loaded
  int m_chunks;                 // 64-column chunks of an MN-major A tile that are loaded
  uint32_t a_tx, b_tx;          // bytes per stage that the loads above deliver
  // bottleneck-isolation switches (sc_debug_set("gemm_dbg", bits); results are WRONG when any bit is set):
  //   1 = epilogue only releases the accumulator, 2 = epilogue without its TMA stores/loads,
  //   4 = no MMAs issued, 8 = every tile loads A block 0 (no HBM reads), 16 = B tiles are not loaded,
  //   32 = no operand loads at all (MMAs run on stale shared memory),
  //   64 = producer / MMA threads hot-spin on their mbarriers (mbarrier.test_wait),
  //   128 = CTA 0 prints the MMA warp's cycles per k-block,
  //   256 = L2 prefetch of upcoming operand tiles (results stay correct; measured slower),
  //   512 = wait accounting (needs a -DSC_GEMM_ACCT build): CTA 0 prints, per role, its total cycles and the cycles
  //         spent in each of its barrier waits
  int dbg;
  int reverse;                  // 1 = items (and the k-blocks of a split-K item) are walked in descending order
  int ln_split;                 // column split: both epilogue warpgroups drain the SAME tile, columns 0-127 / 128-255
                                // (NWG = 2, SLOTS = 2; fused LayerNorm and the output-only staged epilogues)
  int ln_prefetch;              // split LayerNorm epilogue: the NEXT tile's residual chunks are requested into L2 at the start
                                // of this tile (its TMA loads can only be issued once this tile's stores have drained the two
                                // slots, so their latency is exposed at every tile; from L2 it is a third of the HBM latency)
};

struct Item {
  int m_blk, n_blk, kb_begin, kb_end;
};

__device__ __forceinline__ Item decode_item(const GemmParams& p, int item) {
  Item it;
  int s = item % p.ksplit;
  int t = item / p.ksplit;
  it.n_blk = t % p.tiles_n;
  it.m_blk = t / p.tiles_n;
  it.kb_begin = s * p.kb_per_split;
  it.kb_end = min(p.kb_total, it.kb_begin + p.kb_per_split);
  return it;
}

// exact GELU x Phi(x) and its derivative Phi(x) + x phi(x) (SpecV0.gelu == "erf"; generic epilogue paths only)
__device__ __forceinline__ float gelu_erf_fwd(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_bwd(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

template <bool kFast>
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == SC_ACT_GELU) return gelu_fwd<kFast>(x);
  if (act == SC_ACT_GELU_ERF) return gelu_erf_fwd(x);
  if (act == SC_ACT_TANH) return kFast ? tanh_fast(x) : tanhf(x);
  return x;
}

// 8 consecutive columns of one row (vector path: everything 16-byte aligned)
__device__ __forceinline__ void epilogue_vec8(const GemmParams& p, const uint32_t* acc, int64_t row,
                                              int col, uint32_t row_mod, uint32_t row_grp) {
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(acc[j]) * p.alpha;
  if (p.bias) {
    float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
    float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
    x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
    x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
  }
  if (p.rowadd) {
    const float* ra = p.rowadd + static_cast<int64_t>(row_mod) * p.N + col;
    float4 b0 = __ldg(reinterpret_cast<const float4*>(ra));
    float4 b1 = __ldg(reinterpret_cast<const float4*>(ra + 4));
    x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
    x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
  }
  if (p.rowgrp) {
    const float* rg = p.rowgrp + static_cast<int64_t>(row_grp) * p.N + col;
    float4 b0 = __ldg(reinterpret_cast<const float4*>(rg));
    float4 b1 = __ldg(reinterpret_cast<const float4*>(rg + 4));
    const float s = p.rowgrp_scale;
    x[0] = fmaf(b0.x, s, x[0]); x[1] = fmaf(b0.y, s, x[1]); x[2] = fmaf(b0.z, s, x[2]);
    x[3] = fmaf(b0.w, s, x[3]); x[4] = fmaf(b1.x, s, x[4]); x[5] = fmaf(b1.y, s, x[5]);
    x[6] = fmaf(b1.z, s, x[6]); x[7] = fmaf(b1.w, s, x[7]);
  }
  const int64_t off = row * p.ldo + col;
  if (p.aux_out) {
    uint4 w = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]),
                         pack_bf16(x[6], x[7]));
    *reinterpret_cast<uint4*>(p.aux_out + off) = w;
  }
  if (p.act != SC_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = apply_act<true>(x[j], p.act);
  }
  if (p.aux_in) {
    uint4 w = __ldg(reinterpret_cast<const uint4*>(p.aux_in + off));
    float2 a0 = unpack_bf16(w.x), a1 = unpack_bf16(w.y), a2 = unpack_bf16(w.z), a3 = unpack_bf16(w.w);
    const float av[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] *= p.aux_in_erf ? gelu_erf_bwd(av[j]) : gelu_bwd<true>(av[j]);
  }
  if (p.resid) {
    uint4 w = __ldg(reinterpret_cast<const uint4*>(p.resid + off));
    float2 a0 = unpack_bf16(w.x), a1 = unpack_bf16(w.y), a2 = unpack_bf16(w.z), a3 = unpack_bf16(w.w);
    x[0] += a0.x; x[1] += a0.y; x[2] += a1.x; x[3] += a1.y;
    x[4] += a2.x; x[5] += a2.y; x[6] += a3.x; x[7] += a3.y;
  }
  if (p.out_bf16) {
    uint4 w = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]),
                         pack_bf16(x[6], x[7]));
    *reinterpret_cast<uint4*>(p.out_bf16 + off) = w;
  }
  if (p.out_f32) {
    float* o = p.out_f32 + off;
    if (p.atomic) {
      red_add_v4_f32(o, x[0], x[1], x[2], x[3]);
      red_add_v4_f32(o + 4, x[4], x[5], x[6], x[7]);
    } else {
      *reinterpret_cast<float4*>(o) = make_float4(x[0], x[1], x[2], x[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(x[4], x[5], x[6], x[7]);
    }
  }
}

// scalar path for ragged N / unaligned leading dimensions
__device__ __noinline__ void epilogue_scalar(const GemmParams& p, float a, int64_t row, int col,
                                                uint32_t row_mod, uint32_t row_grp) {
  float x = a * p.alpha;
  if (p.bias) x += __ldg(p.bias + col);
  if (p.rowadd) x += __ldg(p.rowadd + static_cast<int64_t>(row_mod) * p.N + col);
  if (p.rowgrp) x = fmaf(__ldg(p.rowgrp + static_cast<int64_t>(row_grp) * p.N + col), p.rowgrp_scale, x);
  const int64_t off = row * p.ldo + col;
  if (p.aux_out) p.aux_out[off] = __float2bfloat16(x);
  x = apply_act<true>(x, p.act);
  if (p.aux_in) {
    const float av = __bfloat162float(p.aux_in[off]);
    x *= p.aux_in_erf ? gelu_erf_bwd(av) : gelu_bwd<true>(av);
  }
  if (p.resid) x += __bfloat162float(p.resid[off]);
  if (p.out_bf16) p.out_bf16[off] = __float2bfloat16(x);
  if (p.out_f32) {
    if (p.atomic) atomicAdd(p.out_f32 + off, x);
    else p.out_f32[off] = x;
  }
}

// ---- packed fp32x2 helpers (FFMA2 / FMUL2 / FADD2: two fp32 lanes per issue slot) ----
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pk2u(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t tanh2(uint64_t u) {
  float a, b;
  upk2(u, a, b);
  return pk2(tanh_fast(a), tanh_fast(b));
}
__device__ __forceinline__ uint32_t pack2_bf16(uint64_t v) {
  float a, b;
  upk2(v, a, b);
  return pack_bf16(a, b);
}
__device__ __forceinline__ uint64_t unpack2_bf16(uint32_t u) {
  return pk2u(u << 16, u & 0xFFFF0000u);
}
// tanh-form GELU on a pair: u = x (c0 + c0 c1 x^2), 0.5 x (1 + tanh u)
__device__ __forceinline__ uint64_t gelu_fwd2(uint64_t x) {
  const uint64_t k0 = pk2(SC_GELU_C0, SC_GELU_C0), k01 = pk2(SC_GELU_C0 * SC_GELU_C1, SC_GELU_C0 * SC_GELU_C1);
  const uint64_t half = pk2(0.5f, 0.5f);
  const uint64_t u = mul2(x, fma2(mul2(x, x), k01, k0));
  const uint64_t hx = mul2(x, half);
  return fma2(hx, tanh2(u), hx);
}
// d/dx of the above: 0.5 (1 + t) + 0.5 x (1 - t^2) c0 (1 + 3 c1 x^2)
__device__ __forceinline__ uint64_t gelu_bwd2(uint64_t x) {
  const uint64_t k0 = pk2(SC_GELU_C0, SC_GELU_C0), k01 = pk2(SC_GELU_C0 * SC_GELU_C1, SC_GELU_C0 * SC_GELU_C1);
  const uint64_t k03 = pk2(3.0f * SC_GELU_C0 * SC_GELU_C1, 3.0f * SC_GELU_C0 * SC_GELU_C1);
  const uint64_t half = pk2(0.5f, 0.5f), one = pk2(1.0f, 1.0f), neg1 = pk2(-1.0f, -1.0f);
  const uint64_t x2 = mul2(x, x);
  const uint64_t t = tanh2(mul2(x, fma2(x2, k01, k0)));
  const uint64_t du = fma2(x2, k03, k0);
  const uint64_t a = add2(t, one);                       // 1 + t
  const uint64_t omt2 = mul2(a, fma2(t, neg1, one));     // (1 + t)(1 - t)
  const uint64_t w = mul2(mul2(x, half), du);
  return fma2(w, omt2, mul2(a, half));
}

// 16 columns of one row for the staged epilogue: math only, results packed to bf16 (8 words out, 8 words of
// pre-activation copy).  KIND != KIND_GEN turns every feature test into a compile-time constant.  `bv` holds
// the 16 bias values of these columns (read from the warpgroup's shared-memory bias tile), `in0/in1` are the 16 bf16 values of the TMA input tile.
template <int KIND>
__device__ __forceinline__ void epilogue_math16(const GemmParams& p, const uint32_t (&acc)[16], const float4 (&bv)[4],
                                                int col, uint32_t row_mod, uint32_t row_grp, const uint4& in0,
                                                const uint4& in1, uint32_t (&outp)[8], uint32_t (&auxp)[8]) {
  constexpr bool G = (KIND == KIND_GEN);
  const bool has_bias = G ? (p.bias != nullptr) : (KIND == KIND_BIAS || KIND == KIND_FC1 || KIND == KIND_RESID);
  const bool want_aux = G ? (p.aux_out != nullptr) : (KIND == KIND_FC1);
  const int act = G ? p.act : (KIND == KIND_FC1 ? SC_ACT_GELU : SC_ACT_NONE);
  const bool in_dgelu = G ? (p.aux_in != nullptr) : (KIND == KIND_DGELU);
  const bool in_resid = G ? (p.epi_mode == EPI_IN && p.aux_in == nullptr) : (KIND == KIND_RESID);
  uint64_t x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = pk2u(acc[2 * j], acc[2 * j + 1]);
  if (G) {
    if (p.alpha != 1.0f) {
      const uint64_t al = pk2(p.alpha, p.alpha);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = mul2(x[j], al);
    }
  }
  if (has_bias) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      x[2 * q] = add2(x[2 * q], pk2(bv[q].x, bv[q].y));
      x[2 * q + 1] = add2(x[2 * q + 1], pk2(bv[q].z, bv[q].w));
    }
  }
  if (G) {
    if (p.rowadd) {
      const float4* ra = reinterpret_cast<const float4*>(p.rowadd + static_cast<int64_t>(row_mod) * p.N + col);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b = __ldg(ra + q);
        x[2 * q] = add2(x[2 * q], pk2(b.x, b.y));
        x[2 * q + 1] = add2(x[2 * q + 1], pk2(b.z, b.w));
      }
    }
    if (p.rowgrp) {
      const float4* rg = reinterpret_cast<const float4*>(p.rowgrp + static_cast<int64_t>(row_grp) * p.N + col);
      const uint64_t sc = pk2(p.rowgrp_scale, p.rowgrp_scale);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b = __ldg(rg + q);
        x[2 * q] = fma2(pk2(b.x, b.y), sc, x[2 * q]);
        x[2 * q + 1] = fma2(pk2(b.z, b.w), sc, x[2 * q + 1]);
      }
    }
  }
  if (want_aux) {
#pragma unroll
    for (int j = 0; j < 8; ++j) auxp[j] = pack2_bf16(x[j]);
  }
  if (act == SC_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = gelu_fwd2(x[j]);
  } else if (G && act == SC_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = tanh2(x[j]);
  }
  if (in_dgelu || in_resid) {
    const uint32_t iw[8] = {in0.x, in0.y, in0.z, in0.w, in1.x, in1.y, in1.z, in1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint64_t a = unpack2_bf16(iw[j]);
      x[j] = in_dgelu ? mul2(x[j], gelu_bwd2(a)) : add2(x[j], a);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) outp[j] = pack2_bf16(x[j]);
}

// Where the NEXT tile of this epilogue warp lives (input-tile prefetch in 4-slot mode).
struct NextTile {
  bool valid;
  int n0, in_row0;
};

// One 128x256 accumulator tile of the staged epilogue for one warp (32 rows): four 64-column chunks, each
// drained in 16-column steps (tcgen05.ld.x16 -> math -> st.shared into a SWIZZLE_128B slot) and then written
// with one (two for EPI_OUT2) TMA store.  The 16 steps are software-pipelined: the tcgen05.ld and the bias
// reads of step s+1 are in flight while step s is computed.
//   SLOTS = 2: chunk c uses slot c & 1 (EPI_OUT2: slot 0 = pre-activation copy, slot 1 = output).  In EPI_IN
//              mode the slot first receives the bf16 input tile by TMA load (chunks 0,1 prefetched before the
//              tile, chunks 2,3 as soon as the store of chunk c-2 has drained the slot) and is overwritten in place.
//   SLOTS = 4: chunk c uses slot c (EPI_OUT2: slots 2(c&1), 2(c&1)+1), so the input of all four chunks of the
//              NEXT tile is requested while this tile is still being drained (slot c-1 is refilled at the end of
//              chunk c) and a two-output chunk never waits for its own stores.
template <int KIND, int SLOTS>
__device__ __forceinline__ void staged_tile(const GemmParams& p, const CUtensorMap* tmOut, const CUtensorMap* tmAux,
                                            const CUtensorMap* tmIn, uint32_t t_addr, uint32_t ring, uint32_t in_bar0,
                                            uint32_t bias_tile, int lane, int n0, int row0, int in_row0,
                                            uint32_t row_mod, uint32_t row_grp, uint32_t tile_parity,
                                            const NextTile& next, int s_lo = 0, int s_hi = 16) {
  // [s_lo, s_hi): the 16-column steps of the tile this warp drains.  All 16, or (column split between two warpgroups
  // that work on the SAME tile, output-only epilogues) steps 0-7 / 8-15 = chunks 0,1 / 2,3.
  constexpr bool G = (KIND == KIND_GEN);
  const bool no_tma = (p.dbg & 2) != 0;
  // The specialised kinds fix the staging mode at compile time (the host only picks a kind whose mode matches), so the
  // hot loop carries no mode branches: the ncu source page showed 74 instructions per 16-column step, 45 of them
  // integer / control overhead, and the epilogue warps issue-bound on them.
  const bool has_in = G ? ((p.epi_mode == EPI_IN) && !no_tma) : ((KIND == KIND_RESID || KIND == KIND_DGELU) && !no_tma);
  const bool two_out = G ? (p.epi_mode == EPI_OUT2) : (KIND == KIND_FC1);
  const bool has_bias = G ? (p.bias != nullptr) : (KIND == KIND_BIAS || KIND == KIND_FC1 || KIND == KIND_RESID);
  const uint32_t sw = static_cast<uint32_t>(lane & 7);
  // swizzled offset of this lane's row: piece (2h | 2h+1) ^ sw of the 128-byte row = one XOR with a constant once h is
  // a compile-time value (the low 7 bits of swz_row are sw << 4)
  const uint32_t swz_row = static_cast<uint32_t>(lane) * 128u | (sw << 4);

  // SLOTS = 1 (output-only epilogues, two warpgroups): one slot per warp, every chunk waits for the store of the chunk before
  // it — the 32 KB that saves buy a FOURTH pipeline stage, i.e. a whole K = 256 tile of operands in flight
  auto out_slot = [&](int c) { return SLOTS == 4 ? (two_out ? 2 * (c & 1) + 1 : c) : (SLOTS == 1 ? 0 : (two_out ? 1 : (c & 1))); };
  auto aux_slot = [&](int c) { return SLOTS == 4 ? 2 * (c & 1) : 0; };

  uint32_t v[2][16];
  float4 bv[2][4];
  auto issue = [&](int s, uint32_t (&vv)[16], float4 (&bb)[4]) {
    tmem_ld_32x32b_x16(t_addr + s * 16, vv);
    if (has_bias) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 w = ld_shared_v4(bias_tile + s * 64 + q * 16);           // same address in every lane: broadcast
        bb[q] = make_float4(__uint_as_float(w.x), __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w));
      }
    }
  };
  auto chunk_begin = [&](int c) {
    if (has_in) {
      if (SLOTS == 4) mbar_wait(in_bar0 + c * 8, tile_parity);               // one use of barrier c per tile
      else mbar_wait(in_bar0 + (c & 1) * 8, static_cast<uint32_t>(c >> 1) & 1u);   // uses 2t + (c >> 1) of barrier c & 1
    } else if (lane == 0) {
      // the TMA store that last read the slot(s) must have drained them
      if (SLOTS == 4) { if (two_out) tma_store_wait_read<1>(); else tma_store_wait_read<3>(); }
      else { if (two_out || SLOTS == 1) tma_store_wait_read<0>(); else tma_store_wait_read<1>(); }
    }
    __syncwarp();
  };
  // one 16-column step: h = its position inside the chunk (compile-time in the unrolled loop below), slot_o / slot_a =
  // shared addresses of the chunk's output / pre-activation slots
  auto process = [&](int s, int h, uint32_t slot_o, uint32_t slot_a, const uint32_t (&vv)[16], const float4 (&bb)[4]) {
    const uint32_t pa = swz_row ^ (32u * h), pb = swz_row ^ (32u * h + 16u);
    uint4 i0 = make_uint4(0u, 0u, 0u, 0u), i1 = i0;
    if (has_in) {
      i0 = ld_shared_v4(slot_o + pa);
      i1 = ld_shared_v4(slot_o + pb);
    }
    uint32_t o[8], a[8];
    epilogue_math16<KIND>(p, vv, bb, n0 + s * 16, row_mod, row_grp, i0, i1, o, a);
    st_shared_v4(slot_o + pa, o[0], o[1], o[2], o[3]);
    st_shared_v4(slot_o + pb, o[4], o[5], o[6], o[7]);
    if (two_out) {
      st_shared_v4(slot_a + pa, a[0], a[1], a[2], a[3]);
      st_shared_v4(slot_a + pb, a[4], a[5], a[6], a[7]);
    }
  };
  auto chunk_end = [&](int c) {
    const int col0 = n0 + c * 64;
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && !no_tma) {
      if (two_out) tma_store_2d(tmAux, ring + aux_slot(c) * EPI_SLOT_BYTES, col0, row0);
      tma_store_2d(tmOut, ring + out_slot(c) * EPI_SLOT_BYTES, col0, row0);
      tma_store_commit();
      if (has_in) {
        if (SLOTS == 4) {
          if (c >= 1 && next.valid) {                                        // slot c-1 -> the next tile's chunk c-1
            tma_store_wait_read<1>();
            mbar_arrive_expect_tx(in_bar0 + (c - 1) * 8, EPI_SLOT_BYTES);
            tma_load_2d(ring + (c - 1) * EPI_SLOT_BYTES, tmIn, in_bar0 + (c - 1) * 8, next.n0 + (c - 1) * 64, next.in_row0);
          }
        } else if (c < 2) {                                                  // refill the slot with chunk c + 2's input
          tma_store_wait_read<0>();
          mbar_arrive_expect_tx(in_bar0 + (c & 1) * 8, EPI_SLOT_BYTES);
          tma_load_2d(ring + (c & 1) * EPI_SLOT_BYTES, tmIn, in_bar0 + (c & 1) * 8, col0 + 128, in_row0);
        }
      }
    }
  };

  // one iteration per 64-column chunk, its four steps unrolled (h is a constant in every process call); the
  // tcgen05.ld of step s + 1 is in flight while step s is computed, across the chunk boundary too
  issue(s_lo, v[0], bv[0]);
#pragma unroll 1
  for (int c = s_lo >> 2; c < (s_hi >> 2); ++c) {
    const int s = c * 4;
    const uint32_t slot_o = ring + out_slot(c) * EPI_SLOT_BYTES;
    const uint32_t slot_a = ring + aux_slot(c) * EPI_SLOT_BYTES;
    chunk_begin(c);
    tmem_ld_wait();
    issue(s + 1, v[1], bv[1]);
    process(s, 0, slot_o, slot_a, v[0], bv[0]);
    tmem_ld_wait();
    issue(s + 2, v[0], bv[0]);
    process(s + 1, 1, slot_o, slot_a, v[1], bv[1]);
    tmem_ld_wait();
    issue(s + 3, v[1], bv[1]);
    process(s + 2, 2, slot_o, slot_a, v[0], bv[0]);
    tmem_ld_wait();
    if (s + 4 < s_hi) issue(s + 4, v[0], bv[0]);
    process(s + 3, 3, slot_o, slot_a, v[1], bv[1]);
    chunk_end(c);
  }
}

// Residual GEMM epilogue with the following LayerNorm fused in (N == 256, 4 slots per warp, 1 warpgroup):
//   pass 1: x = acc + bias + residual tile -> bf16 into the four slots (TMA store, tmOut) while the row statistics
//           accumulate in registers (from the rounded values, like the stand-alone LayerNorm kernel reads them);
//   pass 2: the slots still hold the whole 32 x 256 row block: y = (x - mean) rstd gamma + beta is written over it
//           chunk by chunk once the first store has read the slot, and leaves through tmLn.
// The next tile's residual chunks are requested as the second stores drain.
__device__ __forceinline__ void staged_tile_ln(const GemmParams& p, const CUtensorMap* tmOut, const CUtensorMap* tmLn,
                                               const CUtensorMap* tmIn, uint32_t t_addr, uint32_t ring, uint32_t in_bar0,
                                               uint32_t bias_tile, uint32_t gamma_tile, uint32_t beta_tile, int lane,
                                               int row0, int64_t row, bool row_ok, uint32_t tile_parity,
                                               const NextTile& next) {
  const bool no_tma = (p.dbg & 2) != 0;
  const uint32_t sw = static_cast<uint32_t>(lane & 7);
  auto lds4 = [&](uint32_t addr) {
    const uint4 w = ld_shared_v4(addr);
    return make_float4(__uint_as_float(w.x), __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w));
  };
  uint32_t v[2][16];
  float4 bv[2][4];
  uint64_t s1 = pk2(0.f, 0.f), s2 = s1;
  auto issue = [&](int s, uint32_t (&vv)[16], float4 (&bb)[4]) {
    tmem_ld_32x32b_x16(t_addr + s * 16, vv);
#pragma unroll
    for (int q = 0; q < 4; ++q) bb[q] = lds4(bias_tile + s * 64 + q * 16);
  };
  const uint32_t swz_row = static_cast<uint32_t>(lane) * 128u | (sw << 4);     // see staged_tile
  auto pass1 = [&](int c, int h, const uint32_t (&vv)[16], const float4 (&bb)[4]) {      // h: compile-time in the loop below
    const uint32_t slot = ring + c * EPI_SLOT_BYTES;
    const uint32_t pa = swz_row ^ (32u * h), pb = swz_row ^ (32u * h + 16u);
    uint4 i0 = make_uint4(0u, 0u, 0u, 0u), i1 = i0;
    if (!no_tma) {
      i0 = ld_shared_v4(slot + pa);
      i1 = ld_shared_v4(slot + pb);
    }
    const uint32_t iw[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = bb[j >> 1];
      uint64_t x = add2(pk2u(vv[2 * j], vv[2 * j + 1]), (j & 1) ? pk2(b.z, b.w) : pk2(b.x, b.y));
      x = add2(x, unpack2_bf16(iw[j]));
      o[j] = pack2_bf16(x);
      const uint64_t xr = unpack2_bf16(o[j]);
      s1 = add2(s1, xr);
      s2 = fma2(xr, xr, s2);
    }
    st_shared_v4(slot + pa, o[0], o[1], o[2], o[3]);
    st_shared_v4(slot + pb, o[4], o[5], o[6], o[7]);
  };
  issue(0, v[0], bv[0]);
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    const int s = c * 4;
    if (!no_tma) mbar_wait(in_bar0 + c * 8, tile_parity);
    __syncwarp();
    tmem_ld_wait();
    issue(s + 1, v[1], bv[1]);
    pass1(c, 0, v[0], bv[0]);
    tmem_ld_wait();
    issue(s + 2, v[0], bv[0]);
    pass1(c, 1, v[1], bv[1]);
    tmem_ld_wait();
    issue(s + 3, v[1], bv[1]);
    pass1(c, 2, v[0], bv[0]);
    tmem_ld_wait();
    if (c < 3) issue(s + 4, v[0], bv[0]);
    pass1(c, 3, v[1], bv[1]);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && !no_tma) {
      tma_store_2d(tmOut, ring + c * EPI_SLOT_BYTES, c * 64, row0);
      tma_store_commit();
    }
  }
  float a0, a1, q0, q1;
  upk2(s1, a0, a1);
  upk2(s2, q0, q1);
  const float mean = (a0 + a1) * (1.0f / BN);
  const float var = fmaxf((q0 + q1) * (1.0f / BN) - mean * mean, 0.0f);
  const float rstd = rsqrtf(var + p.ln_eps);
  if (row_ok) {
    if (p.ln_mean) p.ln_mean[row] = mean;
    if (p.ln_rstd) p.ln_rstd[row] = rstd;
  }
  const uint64_t rs2 = pk2(rstd, rstd), nm2 = pk2(-mean * rstd, -mean * rstd);
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    if (lane == 0) tma_store_wait_read<3>();                   // the x store of chunk c has read the slot
    __syncwarp();
    const uint32_t row_o = ring + c * EPI_SLOT_BYTES + lane * 128;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int s = c * 4 + h;
      const uint32_t pa = ((2u * h) ^ sw) << 4, pb = ((2u * h + 1u) ^ sw) << 4;
      const uint4 i0 = ld_shared_v4(row_o + pa), i1 = ld_shared_v4(row_o + pb);
      const uint32_t iw[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
      uint32_t o[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 g = lds4(gamma_tile + s * 64 + q * 16), b = lds4(beta_tile + s * 64 + q * 16);
        o[2 * q] = pack2_bf16(fma2(fma2(unpack2_bf16(iw[2 * q]), rs2, nm2), pk2(g.x, g.y), pk2(b.x, b.y)));
        o[2 * q + 1] = pack2_bf16(fma2(fma2(unpack2_bf16(iw[2 * q + 1]), rs2, nm2), pk2(g.z, g.w), pk2(b.z, b.w)));
      }
      st_shared_v4(row_o + pa, o[0], o[1], o[2], o[3]);
      st_shared_v4(row_o + pb, o[4], o[5], o[6], o[7]);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && !no_tma) {
      tma_store_2d(tmLn, ring + c * EPI_SLOT_BYTES, c * 64, row0);
      tma_store_commit();
      if (c >= 1 && next.valid) {                              // slot c-1 is free: the next tile's residual chunk c-1
        tma_store_wait_read<1>();
        mbar_arrive_expect_tx(in_bar0 + (c - 1) * 8, EPI_SLOT_BYTES);
        tma_load_2d(ring + (c - 1) * EPI_SLOT_BYTES, tmIn, in_bar0 + (c - 1) * 8, next.n0 + (c - 1) * 64, next.in_row0);
      }
    }
  }
}

// Fused LayerNorm with the row split between two epilogue warpgroups (2 slots per warp): warpgroup `half` owns columns
// half*128 .. +127 of every row, i.e. chunks 2*half, 2*half+1 and the 16-column steps 8*half .. +7.  Pass 1 as in
// staged_tile_ln over the own half; the partial (sum x, sum x^2) of a row are exchanged with the partner warp (same TMEM
// lane quarter, other warpgroup) through `xchg` (double-buffered by tile parity) around a 64-thread named barrier; pass 2
// normalises the own half.  One warpgroup needed ~9 us per 128x256 tile (one warp per scheduler, dependency-bound),
// which kept the K = 256 residual GEMMs below their HBM bound.
__device__ __forceinline__ void staged_tile_ln_split(const GemmParams& p, const CUtensorMap* tmOut, const CUtensorMap* tmLn,
                                                     const CUtensorMap* tmIn, uint32_t t_addr, uint32_t ring, uint32_t in_bar0,
                                                     uint32_t bias_tile, uint32_t gamma_tile, uint32_t beta_tile, uint32_t xchg,
                                                     int lane, int half, int quarter, int row0, int64_t row, bool row_ok,
                                                     uint32_t tile_parity, const NextTile& next) {
  const bool no_tma = (p.dbg & 2) != 0;
  const uint32_t sw = static_cast<uint32_t>(lane & 7);
  const int sbase = half * 8, cbase = half * 2;
  auto lds4 = [&](uint32_t addr) {
    const uint4 w = ld_shared_v4(addr);
    return make_float4(__uint_as_float(w.x), __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w));
  };
  uint32_t v[2][16];
  float4 bv[2][4];
  uint64_t s1 = pk2(0.f, 0.f), s2 = s1;
  auto issue = [&](int sl, uint32_t (&vv)[16], float4 (&bb)[4]) {        // sl: local step 0..7
    const int s = sbase + sl;
    tmem_ld_32x32b_x16(t_addr + s * 16, vv);
#pragma unroll
    for (int q = 0; q < 4; ++q) bb[q] = lds4(bias_tile + s * 64 + q * 16);
  };
  const uint32_t swz_row = static_cast<uint32_t>(lane) * 128u | (sw << 4);     // see staged_tile
  auto pass1 = [&](int c, int h, const uint32_t (&vv)[16], const float4 (&bb)[4]) {   // local chunk (slot), group in it
    const uint32_t row_o = ring + c * EPI_SLOT_BYTES;
    const uint32_t pa = swz_row ^ (32u * h), pb = swz_row ^ (32u * h + 16u);
    uint4 i0 = make_uint4(0u, 0u, 0u, 0u), i1 = i0;
    if (!no_tma) {
      i0 = ld_shared_v4(row_o + pa);
      i1 = ld_shared_v4(row_o + pb);
    }
    const uint32_t iw[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = bb[j >> 1];
      uint64_t x = add2(pk2u(vv[2 * j], vv[2 * j + 1]), (j & 1) ? pk2(b.z, b.w) : pk2(b.x, b.y));
      x = add2(x, unpack2_bf16(iw[j]));
      o[j] = pack2_bf16(x);
      const uint64_t xr = unpack2_bf16(o[j]);
      s1 = add2(s1, xr);
      s2 = fma2(xr, xr, s2);
    }
    st_shared_v4(row_o + pa, o[0], o[1], o[2], o[3]);
    st_shared_v4(row_o + pb, o[4], o[5], o[6], o[7]);
  };
  issue(0, v[0], bv[0]);
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    const int sl = c * 4;
    if (!no_tma) mbar_wait(in_bar0 + c * 8, tile_parity);
    __syncwarp();
    tmem_ld_wait();
    issue(sl + 1, v[1], bv[1]);
    pass1(c, 0, v[0], bv[0]);
    tmem_ld_wait();
    issue(sl + 2, v[0], bv[0]);
    pass1(c, 1, v[1], bv[1]);
    tmem_ld_wait();
    issue(sl + 3, v[1], bv[1]);
    pass1(c, 2, v[0], bv[0]);
    tmem_ld_wait();
    if (c < 1) issue(sl + 4, v[0], bv[0]);
    pass1(c, 3, v[1], bv[1]);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && !no_tma) {
      tma_store_2d(tmOut, ring + c * EPI_SLOT_BYTES, (cbase + c) * 64, row0);
      tma_store_commit();
    }
  }
  // exchange the partial row statistics with the warp that owns the other half of these 32 rows
  float a0, a1, q0, q1;
  upk2(s1, a0, a1);
  upk2(s2, q0, q1);
  const uint32_t xb = xchg + tile_parity * (2 * BM * 8);
  const int r = quarter * 32 + lane;
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(xb + (half * BM + r) * 8), "f"(a0 + a1), "f"(q0 + q1) : "memory");
  named_bar_sync(3 + quarter, 64);
  float o0, o1;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(o0), "=f"(o1) : "r"(xb + ((half ^ 1) * BM + r) * 8) : "memory");
  // both halves add in the same order (half 0 first) so the two warps normalise with bit-identical statistics
  const float sum = half == 0 ? (a0 + a1) + o0 : o0 + (a0 + a1);
  const float sq = half == 0 ? (q0 + q1) + o1 : o1 + (q0 + q1);
  const float mean = sum * (1.0f / BN);
  const float var = fmaxf(sq * (1.0f / BN) - mean * mean, 0.0f);
  const float rstd = rsqrtf(var + p.ln_eps);
  if (row_ok && half == 0) {
    if (p.ln_mean) p.ln_mean[row] = mean;
    if (p.ln_rstd) p.ln_rstd[row] = rstd;
  }
  const uint64_t rs2 = pk2(rstd, rstd), nm2 = pk2(-mean * rstd, -mean * rstd);
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    if (lane == 0) tma_store_wait_read<1>();                   // the x store of local chunk c has read the slot
    __syncwarp();
    const uint32_t row_o = ring + c * EPI_SLOT_BYTES + lane * 128;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int s = sbase + c * 4 + h;
      const uint32_t pa = ((2u * h) ^ sw) << 4, pb = ((2u * h + 1u) ^ sw) << 4;
      const uint4 i0 = ld_shared_v4(row_o + pa), i1 = ld_shared_v4(row_o + pb);
      const uint32_t iw[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
      uint32_t o[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 g = lds4(gamma_tile + s * 64 + q * 16), b = lds4(beta_tile + s * 64 + q * 16);
        o[2 * q] = pack2_bf16(fma2(fma2(unpack2_bf16(iw[2 * q]), rs2, nm2), pk2(g.x, g.y), pk2(b.x, b.y)));
        o[2 * q + 1] = pack2_bf16(fma2(fma2(unpack2_bf16(iw[2 * q + 1]), rs2, nm2), pk2(g.z, g.w), pk2(b.z, b.w)));
      }
      st_shared_v4(row_o + pa, o[0], o[1], o[2], o[3]);
      st_shared_v4(row_o + pb, o[4], o[5], o[6], o[7]);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && !no_tma) {
      tma_store_2d(tmLn, ring + c * EPI_SLOT_BYTES, (cbase + c) * 64, row0);
      tma_store_commit();
    }
  }
  // both slots are free once the LayerNorm stores have read them: the next tile's residual chunks of this half
  if (lane == 0 && !no_tma && next.valid) {
    tma_store_wait_read<0>();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      mbar_arrive_expect_tx(in_bar0 + c * 8, EPI_SLOT_BYTES);
      tma_load_2d(ring + c * EPI_SLOT_BYTES, tmIn, in_bar0 + c * 8, next.n0 + (cbase + c) * 64, next.in_row0);
    }
  }
}

template <int A_MN, int B_MN, int CTA2, int NWG, int SLOTS>
__global__ void __launch_bounds__(128 + NWG * 128, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAux,
                    const __grid_constant__ CUtensorMap tmIn, const GemmParams p) {
  using G = Geo<CTA2, NWG, SLOTS>;
  constexpr int STAGES = G::STAGES;
  constexpr int STAGE_BYTES = G::STAGE_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ __align__(8) uint64_t epi_in_bar[MAX_EPI_WARPS][MAX_SLOTS];
  __shared__ uint32_t tmem_base_slot;

  // SWIZZLE_128B atoms need a 1 KB aligned base; the budget in Geo has no slack for re-aligning, so insist on it
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("semcom: dynamic shared memory base 0x%x is not 1 KB aligned\n", smem_base);
    __trap();
  }
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr bool PAIR = (CTA2 == 1);                         // cta_group::2 pair
  constexpr bool MCB = (CTA2 == 2);                          // independent CTAs sharing the B tile by multicast
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;       // 0 = MMA leader of the pair
  const uint32_t crank = CTA2 ? cluster_ctarank() : 0u;      // rank in the cluster (pair or multicast)
  const int ncta = PAIR ? 2 : 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), MCB ? 2 : 1);     // MCB: both CTAs' MMA warps release a shared stage
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tmem_full_bar[s]), 1);
      // pair: both CTAs' epilogue warpgroups arrive on the leader's; split LayerNorm: both warpgroups drain every tile
      mbar_init(smem_u32(&tmem_empty_bar[s]), 128 * ncta * (p.ln_split ? 2 : 1));
    }
    for (int w = 0; w < MAX_EPI_WARPS; ++w)
      for (int s = 0; s < MAX_SLOTS; ++s) mbar_init(smem_u32(&epi_in_bar[w][s]), 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_2sm<TMEM_COLS>(smem_u32(&tmem_base_slot));
    else tmem_alloc<TMEM_COLS>(smem_u32(&tmem_base_slot));
  }
  // bias-gradient mode: 4 KB of bf16 ones in the (otherwise unused) epilogue ring = the B operand of the
  // extra row-sum MMA; its layout is irrelevant because every element is 1.0
  const uint32_t ones_tile = smem_base + G::SMEM;
  if (p.colsum) {
    for (int i = threadIdx.x; i < 4096 / 16; i += blockDim.x)
      st_shared_v4(ones_tile + i * 16, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
  }
  // everything above touches only shared memory / TMEM / kernel parameters: with programmatic dependent launch it
  // overlaps the tail of the preceding kernel.  Global memory is read or written only after this point.
  pdl_sync();
  if (G::LN_TILES && p.ln_gamma) {                           // fused LayerNorm: gamma / beta of the 256 columns
    const uint32_t ln_base = smem_base + G::SMEM + G::RING + NWG * BIAS_TILE_BYTES;
    for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) {
      const float w = i < BN ? __ldg(p.ln_gamma + i) : __ldg(p.ln_beta + i - BN);
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(ln_base + i * 4), "f"(w) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();            // peer barriers are initialised before anyone signals them (pair and MCB)
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // CTA2: p.tiles_m counts 256-row tile pairs; this CTA owns 128-row block 2 * pair + rank of each
  const int num_items = p.tiles_m * p.tiles_n * p.ksplit;
  const int cta_stride = CTA2 ? static_cast<int>(gridDim.x) / 2 : static_cast<int>(gridDim.x);
  const int cta_first = CTA2 ? static_cast<int>(blockIdx.x) / 2 : static_cast<int>(blockIdx.x);
  auto get_item = [&](int item) {
    Item it = decode_item(p, p.reverse ? num_items - 1 - item : item);
    if (CTA2) it.m_blk = it.m_blk * 2 + static_cast<int>(crank);
    return it;
  };
  const bool single_acc = (p.colsum != nullptr);       // stage 1's columns hold the row-sum accumulator

  if (warp == 0) {
    // ===================== TMA producer =====================
    // All 32 lanes run the loop converged, one elected lane issues (see elect_one_sync); the test of the NEXT
    // stage's empty barrier is issued before this stage's loads so its latency hides behind them.
    {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t ready = mbar_test(smem_u32(&empty_bar[0]), 1u);
      const bool acct = kAcct && (p.dbg & 512) && blockIdx.x == 0;
      const long long acct_t0 = clock64();
      long long acct_wait = 0;
      for (int item = cta_first; item < num_items; item += cta_stride) {
        const Item it = get_item(item);
        const int m0 = (p.dbg & 8) ? 0 : it.m_blk * BM;
        const int n0 = it.n_blk * BN + static_cast<int>(rank) * G::B_ROWS;   // pair: this CTA's half of B
        // L2 prefetch (opt-in: gemm_dbg bit 256).  Activation-times-weight GEMMs: the A tile of this CTA's NEXT item
        // is requested a whole tile ahead; split-K wgrad: both operands kPrefetchKb k-blocks ahead.  Measured NEGATIVE
        // (profiles/README.md, r1 "L2 prefetch"): the whole step got 11 % slower, K = 1024 shapes and wgrads by
        // 25-35 % -- the memory system, not the ring depth, is what these kernels wait for.
        const bool l2_pf = !CTA2 && (p.dbg & 256);
        if (l2_pf && p.ksplit == 1 && elect_one_sync()) {
          const int nitem = item + cta_stride;
          if (nitem < num_items) {
            const Item nx = get_item(nitem);
            const int nm0 = nx.m_blk * BM;
            const int nkb = min(nx.kb_end - nx.kb_begin, 16);
            for (int kb = 0; kb < nkb; ++kb) {
              if (A_MN == 0) {
                tma_prefetch_l2_2d(&tmA, (nx.kb_begin + kb) * BK, nm0);
              } else {
                for (int c = 0; c < p.m_chunks; ++c) tma_prefetch_l2_2d(&tmA, nm0 + 64 * c, (nx.kb_begin + kb) * BK);
              }
            }
          }
        }
        constexpr int kPrefetchKb = 6;
        for (int kb = it.kb_begin; kb < it.kb_end; ++kb) {
          if (l2_pf && p.ksplit > 1 && kb + kPrefetchKb < it.kb_end && elect_one_sync()) {
            const int pk0 = (kb + kPrefetchKb) * BK;
            if (A_MN == 0) {
              tma_prefetch_l2_2d(&tmA, pk0, m0);
            } else {
              for (int c = 0; c < p.m_chunks; ++c) tma_prefetch_l2_2d(&tmA, m0 + 64 * c, pk0);
            }
            if (B_MN == 0) {
              tma_prefetch_l2_2d(&tmB, pk0, n0);
            } else {
              for (int c = 0; c < p.n_chunks; ++c) tma_prefetch_l2_2d(&tmB, n0 + 64 * c, pk0);
            }
          }
          const uint32_t fb_local = smem_u32(&full_bar[stage]);
          if (!ready) {
            const long long tw = acct ? clock64() : 0;
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
            if (acct) acct_wait += clock64() - tw;
          }
          const int nstage = (stage + 1 == STAGES) ? 0 : stage + 1;
          const uint32_t nphase = (stage + 1 == STAGES) ? (phase ^ 1u) : phase;
          ready = mbar_test(smem_u32(&empty_bar[nstage]), nphase ^ 1u);
          const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
          const uint32_t b_dst = a_dst + A_BYTES;
          const int k0 = (p.reverse ? it.kb_end - 1 - (kb - it.kb_begin) : kb) * BK;
          if (elect_one_sync()) {
            if (p.dbg & 32) {                                   // no loads at all: raw MMA issue rate on stale smem
              if (!PAIR || rank == 0) mbar_arrive(fb_local);
            } else if (MCB) {
              // own A tile (unicast) + this CTA's half of the B tile, multicast into both CTAs; each CTA's full barrier
              // expects its A bytes and the whole B tile (two halves from two issuers)
              const uint32_t fb = fb_local;
              mbar_arrive_expect_tx(fb, p.a_tx + p.b_tx);
              if (A_MN == 0) {
                tma_load_2d(a_dst, &tmA, fb, k0, m0);
              } else {
                for (int c = 0; c < p.m_chunks; ++c) tma_load_2d(a_dst + c * CHUNK_BYTES, &tmA, fb, m0 + 64 * c, k0);
              }
              if (B_MN == 0) {                                  // K-major: rows crank*128 .. +127 of the tile (16 KB)
                tma_load_2d_mc(b_dst + crank * (BN / 2) * BK * 2, &tmB, fb, k0, n0 + static_cast<int>(crank) * (BN / 2), 3);
              } else {                                          // MN-major: 64-column chunks 2*crank, 2*crank + 1
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                  const int cc = 2 * static_cast<int>(crank) + c;
                  tma_load_2d_mc(b_dst + cc * CHUNK_BYTES, &tmB, fb, n0 + 64 * cc, k0, 3);
                }
              }
            } else if (PAIR) {
              // all bytes of both CTAs are accounted on the LEADER's full barrier
              const uint32_t fb = mapa_shared(fb_local, 0);
              if (rank == 0) mbar_arrive_expect_tx(fb_local, 2 * STAGE_BYTES);
              if (A_MN == 0) {
                tma_load_2d_2sm(a_dst, &tmA, fb, k0, m0);
              } else {
#pragma unroll
                for (int c = 0; c < BM / 64; ++c) tma_load_2d_2sm(a_dst + c * CHUNK_BYTES, &tmA, fb, m0 + 64 * c, k0);
              }
              if (B_MN == 0) {
                tma_load_2d_2sm(b_dst, &tmB, fb, k0, n0);
              } else {
#pragma unroll
                for (int c = 0; c < G::B_ROWS / 64; ++c)
                  tma_load_2d_2sm(b_dst + c * CHUNK_BYTES, &tmB, fb, n0 + 64 * c, k0);
              }
            } else {
              const uint32_t fb = fb_local;
              mbar_arrive_expect_tx(fb, (p.dbg & 16) ? p.a_tx : p.a_tx + p.b_tx);
              if (A_MN == 0) {
                tma_load_2d(a_dst, &tmA, fb, k0, m0);
              } else {
                for (int c = 0; c < p.m_chunks; ++c) tma_load_2d(a_dst + c * CHUNK_BYTES, &tmA, fb, m0 + 64 * c, k0);
              }
              if (p.dbg & 16) {
              } else if (B_MN == 0) {
                tma_load_2d(b_dst, &tmB, fb, k0, n0);                      // box = 64 k x n_mma rows
              } else {
                for (int c = 0; c < p.n_chunks; ++c) tma_load_2d(b_dst + c * CHUNK_BYTES, &tmB, fb, n0 + 64 * c, k0);
              }
            }
          }
          __syncwarp();
          stage = nstage;
          phase = nphase;
        }
      }
      if (acct && lane == 0)
        printf("semcom gemm acct: producer  total %lld clk, waiting for a free stage %lld\n", clock64() - acct_t0, acct_wait);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // All 32 lanes run this loop converged (warp-uniform control flow and operands); one elected lane issues the
    // tcgen05 instructions.  Descriptors are a per-stage constant plus k * step in the 14-bit start-address field.
    // Per k-block the issuing lane needs ~4 x 60 cycles for the MMAs + ~65 for the commit; a blocking wait in front
    // (~170 cycles even when the phase is complete) would push that past the 512 cycles the tensor pipe needs, so
    // the NEXT stage's full barrier is tested before this stage's MMAs are issued and only waited on if not ready.
    if (rank == 0) {                                           // pair: only the leader CTA issues MMAs
      const uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * BM : BM, PAIR ? BN : static_cast<uint32_t>(p.n_mma), A_MN, B_MN);
      const uint64_t a_desc0 = umma_smem_desc(0, p.a_lbo, p.a_sbo);
      const uint64_t b_desc0 = umma_smem_desc(0, p.b_lbo, p.b_sbo);
      const uint32_t a_step = p.a_kstep >> 4, b_step = p.b_kstep >> 4;
      const bool mma_off = (p.dbg & 4) != 0;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const long long t_begin = clock64();
      long long n_kb = 0;
      const bool acct = kAcct && (p.dbg & 512) && blockIdx.x == 0;
      long long acct_acc = 0, acct_full = 0;
      uint32_t ready = mbar_test(smem_u32(&full_bar[0]), 0u);
      for (int item = cta_first; item < num_items; item += cta_stride) {
        const Item it = get_item(item);
        {
          const long long tw = acct ? clock64() : 0;
          mbar_wait(smem_u32(&tmem_empty_bar[as]), aphase ^ 1u);
          if (acct) acct_acc += clock64() - tw;
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        const bool rowsum = single_acc && it.n_blk == 0;
        n_kb += it.kb_end - it.kb_begin;
        for (int kb = it.kb_begin; kb < it.kb_end; ++kb) {
          if (!ready) {
            const long long tw = acct ? clock64() : 0;
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            if (acct) acct_full += clock64() - tw;
          }
          const int nstage = (stage + 1 == STAGES) ? 0 : stage + 1;
          const uint32_t nphase = (stage + 1 == STAGES) ? (phase ^ 1u) : phase;
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * STAGE_BYTES;
          // 18-bit CTA-local address: in a cluster the shared::cta window of rank r starts at r << 24 (would spill into LBO)
          const uint64_t ad = a_desc0 + ((a_addr & 0x3FFFFu) >> 4);
          const uint64_t bd = b_desc0 + (((a_addr + A_BYTES) & 0x3FFFFu) >> 4);
          const uint32_t first = (kb > it.kb_begin) ? 1u : 0u;
          if (!PAIR && !rowsum && !mma_off) {
            // common case: test of the next stage, four MMAs and the commit in one asm block (see common.cuh);
            // MCB: the commit arrives on the stage's empty barrier in BOTH CTAs
            uint32_t r = 0;
            if (elect_one_sync())
              r = umma_kblock4_bf16_ss<true, true, MCB>(d_tmem, ad, bd, a_step, b_step, idesc, first,
                                                        smem_u32(&empty_bar[stage]), smem_u32(&full_bar[nstage]), nphase);
            ready = __shfl_sync(0xffffffffu, r, 0);            // elect.sync picks lane 0 of the converged warp
          } else {
            ready = mbar_test(smem_u32(&full_bar[nstage]), nphase);
            if (elect_one_sync()) {
              if (!mma_off) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  const uint32_t acc = (k > 0) ? 1u : first;
                  if (PAIR) umma_bf16_ss_2sm(d_tmem, ad + k * a_step, bd + k * b_step, idesc, acc);
                  else umma_bf16_ss(d_tmem, ad + k * a_step, bd + k * b_step, idesc, acc);
                  if (rowsum) {
                    constexpr uint32_t idesc_ones = umma_idesc_bf16(PAIR ? 2 * BM : BM, 16, A_MN, 1);
                    const uint64_t od = umma_smem_desc(ones_tile, CHUNK_BYTES, 1024);
                    if (PAIR) umma_bf16_ss_2sm(tmem_base + BN, ad + k * a_step, od, idesc_ones, acc);
                    else umma_bf16_ss(tmem_base + BN, ad + k * a_step, od, idesc_ones, acc);
                  }
                }
              }
              // smem slot reusable (in both CTAs of a pair) when the MMAs retire
              if (PAIR) umma_commit_2sm(smem_u32(&empty_bar[stage]));
              else if (MCB) umma_commit_mc2(smem_u32(&empty_bar[stage]));
              else umma_commit(smem_u32(&empty_bar[stage]));
            }
            __syncwarp();
          }
          stage = nstage;
          phase = nphase;
        }
        // accumulator complete (published to both CTAs of a pair)
        if (elect_one_sync()) {
          if (PAIR) umma_commit_2sm(smem_u32(&tmem_full_bar[as])); else umma_commit(smem_u32(&tmem_full_bar[as]));
        }
        __syncwarp();
        if (single_acc) {
          aphase ^= 1u;
        } else {
          as ^= 1;
          if (as == 0) aphase ^= 1u;
        }
      }
      if (acct && lane == 0)
        printf("semcom gemm acct: mma       total %lld clk, %lld k-blocks, waiting for a free accumulator %lld, for operands %lld\n",
               clock64() - t_begin, n_kb, acct_acc, acct_full);
      if ((p.dbg & 128) && blockIdx.x == 0 && lane == 0)       // issue-loop pace: 512 clk per k-block is the tensor pipe's floor
        printf("semcom gemm: MMA warp issued %lld k-blocks in %lld clk (%.0f clk per k-block)\n", n_kb,
               clock64() - t_begin, double(clock64() - t_begin) / double(n_kb > 0 ? n_kb : 1));
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // two warpgroups: warps 4-7 drain accumulator stage 0 (even tiles), warps 8-11 stage 1 (odd tiles)
    // NWG = 2: warps 4-7 take the even tiles of this CTA, warps 8-11 the odd ones; NWG = 1: warps 4-7 take all.
    // Tile n of the CTA always uses accumulator stage n & 1 and phase (n >> 1) & 1.
    const int ew = warp - 4;                                   // 0 .. 4*NWG-1
    const int wg = ew >> 2;
    const int q = warp & 3;                                    // TMEM lane quarter this warp may access
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // column split: the two warpgroups work on the SAME tile (columns 0-127 / 128-255) instead of alternate tiles.
    // With alternate tiles a warpgroup idles while the MMAs refill ITS accumulator (epilogue -> MMA -> epilogue on one
    // buffer); as one double-speed consumer of double-buffered accumulators the MMAs of tile i+1 run under the drain of
    // tile i (ncu: the epilogue warps of the qkv GEMM spent 35 % of their samples waiting for accumulators).
    const bool ln_split = (NWG == 2 && SLOTS == 2 && CTA2 != 1) && p.ln_split != 0;
    const int stride2 = ln_split ? cta_stride : NWG * cta_stride;
    const int first = ln_split ? cta_first : cta_first + wg * cta_stride;
    // pair: the leader's MMA thread waits for BOTH CTAs' epilogues -> arrive on the leader's barrier
    auto release_acc = [&](int as) {
      const uint32_t local = smem_u32(&tmem_empty_bar[as]);
      if (PAIR) mbar_arrive_cluster(mapa_shared(local, 0)); else mbar_arrive(local);
    };
    if (p.epi_mode == EPI_GENERIC) {
      const int g_first = single_acc ? cta_first : first;
      const int g_stride = single_acc ? cta_stride : stride2;
      int n = single_acc ? 0 : wg;                             // index of the tile within this CTA's sequence
      const int n_step = single_acc ? 1 : NWG;
      for (int item = (single_acc && wg == 1) ? num_items : g_first; item < num_items; item += g_stride, n += n_step) {
        const Item it = get_item(item);
        const int as = single_acc ? 0 : (n & 1);
        const uint32_t par = single_acc ? static_cast<uint32_t>(n & 1) : static_cast<uint32_t>((n >> 1) & 1);
        const uint32_t t_addr = t_lane + static_cast<uint32_t>(as * BN);
        const int64_t row = static_cast<int64_t>(it.m_blk) * BM + q * 32 + lane;
        const int n0 = it.n_blk * BN;
        const bool row_ok = row < p.M;
        uint32_t row_mod = 0, row_grp = 0;
        if (p.rowadd) row_mod = static_cast<uint32_t>(row) % p.rowadd_period;
        if (p.rowgrp) row_grp = static_cast<uint32_t>(row) / p.rowgrp_size;
        mbar_wait(smem_u32(&tmem_full_bar[as]), par);
        __syncwarp();
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int col0 = n0 + c * 32;
          if (col0 >= p.N) break;
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_addr + c * 32, v);
          tmem_ld_wait();
          if (row_ok) {
            if (p.vec_ok) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const int col = col0 + g * 8;
                if (col < p.N) epilogue_vec8(p, &v[g * 8], row, col, row_mod, row_grp);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int col = col0 + j;
                if (col < p.N) epilogue_scalar(p, __uint_as_float(v[j]), row, col, row_mod, row_grp);
              }
            }
          }
        }
        if (single_acc && it.n_blk == 0) {
          uint32_t rs;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(rs) : "r"(t_addr + BN) : "memory");
          tmem_ld_wait();
          if (row_ok) atomicAdd(p.colsum + row, __uint_as_float(rs));
        }
        tc_fence_before();
        release_acc(as);
      }
    } else {
      // ---- staged epilogue: 64-column chunks through per-warp SWIZZLE_128B slots, outputs by TMA store
      //      (box 64 x 32), the optional bf16 input tile (residual / GELU' operand / pos table) by TMA load ----
      const uint32_t ring = smem_base + G::SMEM + ew * (SLOTS * EPI_SLOT_BYTES);
      const uint32_t bias_tile = smem_base + G::SMEM + G::RING + wg * BIAS_TILE_BYTES;
      const int wg_tid = static_cast<int>(threadIdx.x) - 128 - wg * 128;      // 0..127 within this warpgroup
      const bool has_in = (p.epi_mode == EPI_IN) && !(p.dbg & 3);
      const bool stage_bias = (p.bias != nullptr) && !(p.dbg & 1);
      const uint32_t ib0 = smem_u32(&epi_in_bar[ew][0]);
      auto in_row = [&](int row0) { return p.in_row_period > 0 ? row0 % p.in_row_period : row0; };
      constexpr int PRE = SLOTS == 4 ? 4 : (SLOTS == 1 ? 0 : 2);               // input chunks requested ahead of a tile
      if (has_in && lane == 0 && first < num_items) {
        const Item it = get_item(first);
#pragma unroll
        for (int c = 0; c < PRE; ++c) {
          mbar_arrive_expect_tx(ib0 + c * 8, EPI_SLOT_BYTES);
          tma_load_2d(ring + c * EPI_SLOT_BYTES, &tmIn, ib0 + c * 8, it.n_blk * BN + c * 64 + (ln_split ? wg * 128 : 0),
                      in_row(it.m_blk * BM + q * 32));
        }
      }
      int n = ln_split ? 0 : wg;
      uint32_t tile_parity = 0;
      const bool acct = kAcct && (p.dbg & 512) && blockIdx.x == 0 && (warp & 3) == 0;
      const long long acct_t0 = clock64();
      long long acct_wait = 0;
      int acct_tiles = 0;
      for (int item = first; item < num_items; item += stride2, n += (ln_split ? 1 : NWG), tile_parity ^= 1u) {
        const Item it = get_item(item);
        const int as = n & 1;
        const uint32_t t_addr = t_lane + static_cast<uint32_t>(as * BN);
        const int row0 = it.m_blk * BM + q * 32;
        const int64_t row = min(static_cast<int64_t>(row0) + lane, p.M - 1);   // tail rows: TMA clips the store
        const int n0 = it.n_blk * BN;
        uint32_t row_mod = 0, row_grp = 0;
        if (p.rowadd) row_mod = static_cast<uint32_t>(row) % p.rowadd_period;
        if (p.rowgrp) row_grp = static_cast<uint32_t>(row) / p.rowgrp_size;
        // this tile's 256 bias values: fetched before the accumulator wait, parked in the warpgroup's smem tile
        float2 bias2 = make_float2(0.f, 0.f);
        if (stage_bias) bias2 = __ldg(reinterpret_cast<const float2*>(p.bias + n0) + wg_tid);
        {
          const long long tw = acct ? clock64() : 0;
          mbar_wait(smem_u32(&tmem_full_bar[as]), static_cast<uint32_t>((n >> 1) & 1));
          if (acct) { acct_wait += clock64() - tw; ++acct_tiles; }
        }
        __syncwarp();
        tc_fence_after();
        const int ir0 = in_row(row0);
        const int next_item = item + stride2;
        NextTile next;
        next.valid = next_item < num_items;
        next.n0 = 0; next.in_row0 = 0;
        if (next.valid) {
          const Item nt = get_item(next_item);
          next.n0 = nt.n_blk * BN;
          next.in_row0 = in_row(nt.m_blk * BM + q * 32);
        }
        if (p.dbg & 1) {
          tc_fence_before();
          release_acc(as);
          continue;
        }
        if (stage_bias) {
          named_bar_sync(1 + wg, 128);                                         // previous tile's bias reads are done
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(bias_tile + wg_tid * 8), "f"(bias2.x), "f"(bias2.y) : "memory");
          named_bar_sync(1 + wg, 128);
        }
        switch (p.epi_kind) {
#define SC_TILE(KIND) staged_tile<KIND, SLOTS>(p, &tmOut, &tmAux, &tmIn, t_addr, ring, ib0, bias_tile, lane, n0, row0, ir0, row_mod, row_grp, tile_parity, next, ln_split ? wg * 8 : 0, ln_split ? wg * 8 + 8 : 16)
          case KIND_PLAIN: SC_TILE(KIND_PLAIN); break;
          case KIND_BIAS:  SC_TILE(KIND_BIAS); break;
          case KIND_FC1:   SC_TILE(KIND_FC1); break;
          case KIND_RESID: SC_TILE(KIND_RESID); break;
          case KIND_DGELU: SC_TILE(KIND_DGELU); break;
          case KIND_RESID_LN:
            if constexpr (G::LN_TILES != 0) {
              const uint32_t ln_base = smem_base + G::SMEM + G::RING + NWG * BIAS_TILE_BYTES;
              if constexpr (G::LN_XCHG != 0) {
                if (ln_split && p.ln_prefetch && lane == 0 && next.valid && p.in_row_period == 0 && !(p.dbg & 3)) {
                  tma_prefetch_l2_2d(&tmIn, next.n0 + (wg * 2) * 64, next.in_row0);
                  tma_prefetch_l2_2d(&tmIn, next.n0 + (wg * 2 + 1) * 64, next.in_row0);
                }
                if (ln_split)
                  staged_tile_ln_split(p, &tmOut, &tmAux, &tmIn, t_addr, ring, ib0, bias_tile, ln_base, ln_base + BIAS_TILE_BYTES,
                                       ln_base + 2 * BIAS_TILE_BYTES, lane, wg, q, row0, static_cast<int64_t>(row0) + lane,
                                       static_cast<int64_t>(row0) + lane < p.M, tile_parity, next);
              } else {
                staged_tile_ln(p, &tmOut, &tmAux, &tmIn, t_addr, ring, ib0, bias_tile, ln_base, ln_base + BIAS_TILE_BYTES, lane,
                               row0, static_cast<int64_t>(row0) + lane, static_cast<int64_t>(row0) + lane < p.M, tile_parity, next);
              }
            }
            break;
          default:         SC_TILE(KIND_GEN); break;
#undef SC_TILE
        }
        tc_fence_before();
        release_acc(as);
        // input chunks of this warp's next tile that could not be requested inside the tile: the slots are free once
        // the stores that read them have drained
        if (has_in && lane == 0 && next.valid && !ln_split) {
          tma_store_wait_read<0>();
#pragma unroll
          for (int c = (SLOTS == 4 ? 3 : 0); c < PRE; ++c) {
            mbar_arrive_expect_tx(ib0 + c * 8, EPI_SLOT_BYTES);
            tma_load_2d(ring + c * EPI_SLOT_BYTES, &tmIn, ib0 + c * 8, next.n0 + c * 64, next.in_row0);
          }
        }
      }
      if (lane == 0) tma_store_wait_all();
      if (acct && lane == 0)
        printf("semcom gemm acct: epilogue warpgroup %d  total %lld clk, %d tiles, waiting for accumulators %lld\n", wg,
               clock64() - acct_t0, acct_tiles, acct_wait);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();            // nobody tears down while the peer may still signal / read it
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2sm<TMEM_COLS>(tmem_base); else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

struct DebugOverrides {
  int64_t a_lbo = 0, a_sbo = 0, a_kstep = 0, b_lbo = 0, b_sbo = 0, b_kstep = 0, max_ctas = 0, generic_epi = 0, cta2 = 0, nwg = 0,
          gemm_dbg = 0, slots = 0, cta2_wgrad = 1, cta2_mink = 512, ln_split = 1, mcb = 0, col_split = 0, ln_prefetch = 0, one_slot = 1;
};
DebugOverrides g_dbg;

template <int A_MN, int B_MN, int CTA2, int NWG, int SLOTS>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmAux,
           const CUtensorMap& tmIn, const GemmParams& p, int grid, cudaStream_t st) {
  static bool attr_set_dev[64] = {};                         // the opt-in is a per-device attribute of the function
  const int dev = sc_device_index();
  if (dev < 0) { sc_set_error("sc_gemm_bf16: cudaGetDevice failed"); return SC_E_LAUNCH; }
  bool& attr_set = attr_set_dev[dev];
  auto kern = gemm_tcgen05_kernel<A_MN, B_MN, CTA2, NWG, SLOTS>;
  constexpr int smem = Geo<CTA2, NWG, SLOTS>::TOTAL;
  constexpr int threads = Geo<CTA2, NWG, SLOTS>::THREADS;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      sc_set_error("sc_gemm_bf16: cannot set %d B dynamic smem: %s", smem, cudaGetErrorString(e));
      return SC_E_LAUNCH;
    }
    attr_set = true;
  }
  if (CTA2) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOut, tmAux, tmIn, p);
    if (e != cudaSuccess) {
      sc_set_error("sc_gemm_bf16: cluster launch failed: %s", cudaGetErrorString(e));
      return SC_E_LAUNCH;
    }
  } else {
    cudaError_t e = sc_launch_kernel(kern, dim3(grid), dim3(threads), smem, st, tmA, tmB, tmOut, tmAux, tmIn, p);
    if (e != cudaSuccess) {
      sc_set_error("sc_gemm_bf16: launch failed: %s", cudaGetErrorString(e));
      return SC_E_LAUNCH;
    }
  }
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_gemm_bf16");
  return SC_OK;
}

}  // namespace

void sc_attn_debug_set_fwd_stages(int v);
void sc_pdl_set(int v);
void sc_mlp_debug_set(int v);
void sc_alternate_set(int v);
void sc_tmap_cache_set(int on);
extern "C" int g_chan_exact;
extern "C" int g_chan_generic;

extern "C" int sc_debug_set(const char* key, int64_t value) {
  if (!key) return SC_E_SHAPE;
  if (!strcmp(key, "attn_fwd_stages")) { sc_attn_debug_set_fwd_stages((int)value); return SC_OK; }
  if (!strcmp(key, "pdl")) { sc_pdl_set((int)value); return SC_OK; }
  if (!strcmp(key, "mlp_dbg")) { sc_mlp_debug_set((int)value); return SC_OK; }
  if (!strcmp(key, "alternate")) { sc_alternate_set((int)value); return SC_OK; }
  if (!strcmp(key, "tmap_cache")) { sc_tmap_cache_set((int)value); return SC_OK; }
  if (!strcmp(key, "chan_exact")) { g_chan_exact = (int)value; return SC_OK; }
  if (!strcmp(key, "chan_generic")) { g_chan_generic = (int)value; return SC_OK; }
  if (!strcmp(key, "a_lbo")) g_dbg.a_lbo = value;
  else if (!strcmp(key, "a_sbo")) g_dbg.a_sbo = value;
  else if (!strcmp(key, "a_kstep")) g_dbg.a_kstep = value;
  else if (!strcmp(key, "b_lbo")) g_dbg.b_lbo = value;
  else if (!strcmp(key, "b_sbo")) g_dbg.b_sbo = value;
  else if (!strcmp(key, "b_kstep")) g_dbg.b_kstep = value;
  else if (!strcmp(key, "max_ctas")) g_dbg.max_ctas = value;
  else if (!strcmp(key, "generic_epi")) g_dbg.generic_epi = value;
  else if (!strcmp(key, "cta2")) g_dbg.cta2 = value;
  else if (!strcmp(key, "cta2_wgrad")) g_dbg.cta2_wgrad = value;
  else if (!strcmp(key, "cta2_mink")) g_dbg.cta2_mink = value;
  else if (!strcmp(key, "ln_split")) g_dbg.ln_split = value;
  else if (!strcmp(key, "mcb")) g_dbg.mcb = value;
  else if (!strcmp(key, "col_split")) g_dbg.col_split = value;
  else if (!strcmp(key, "ln_prefetch")) g_dbg.ln_prefetch = value;
  else if (!strcmp(key, "one_slot")) g_dbg.one_slot = value;
  else if (!strcmp(key, "nwg")) g_dbg.nwg = value;
  else if (!strcmp(key, "gemm_dbg")) g_dbg.gemm_dbg = value;
  else if (!strcmp(key, "slots")) g_dbg.slots = value;
  else if (!strcmp(key, "reset")) g_dbg = DebugOverrides();
  else {
    sc_set_error("sc_debug_set: unknown key %s", key);
    return SC_E_SHAPE;
  }
  return SC_OK;
}

extern "C" int sc_gemm_bf16(const void* A, const void* B, int64_t M, int64_t N, int64_t K, int a_mn,
                            int b_mn, int64_t lda, int64_t ldb, const sc_gemm_epilogue* epi, int ksplit,
                            void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(A && B && epi, SC_E_SHAPE, "sc_gemm_bf16: null operand");
  SC_REQUIRE(M > 0 && N > 0 && K > 0, SC_E_SHAPE, "sc_gemm_bf16: empty problem M=%lld N=%lld K=%lld",
             (long long)M, (long long)N, (long long)K);
  SC_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), SC_E_SHAPE, "sc_gemm_bf16: dims must fit int32");
  SC_REQUIRE((a_mn == 0 || a_mn == 1) && (b_mn == 0 || b_mn == 1), SC_E_SHAPE, "sc_gemm_bf16: a_mn/b_mn must be 0/1");
  SC_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, SC_E_ALIGN, "sc_gemm_bf16: lda=%lld ldb=%lld must be multiples of 8",
             (long long)lda, (long long)ldb);
  SC_REQUIRE(lda >= (a_mn ? M : K) && ldb >= (b_mn ? N : K), SC_E_SHAPE, "sc_gemm_bf16: leading dimension too small");
  SC_REQUIRE(sc_aligned16(A) && sc_aligned16(B), SC_E_ALIGN, "sc_gemm_bf16: A/B must be 16-byte aligned");
  SC_REQUIRE(epi->out_bf16 || epi->out_f32, SC_E_SHAPE, "sc_gemm_bf16: no output pointer");
  SC_REQUIRE(epi->ldo >= N, SC_E_SHAPE, "sc_gemm_bf16: ldo < N");
  SC_REQUIRE(!epi->rowadd || epi->rowadd_period > 0, SC_E_SHAPE, "sc_gemm_bf16: rowadd_period must be > 0");
  SC_REQUIRE(!epi->rowgrp || epi->rowgrp_size > 0, SC_E_SHAPE, "sc_gemm_bf16: rowgrp_size must be > 0");
  SC_REQUIRE(epi->act >= SC_ACT_NONE && epi->act <= SC_ACT_GELU_ERF, SC_E_SHAPE, "sc_gemm_bf16: bad act");

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  // cta2_wgrad: CTA pairs (cta_group::2, 256 x 256 tile per pair, each CTA loads half of B) for the split-K weight
  // gradients only: one long item per pair, so the accumulator hand-off that hurt the K = 256 shapes happens once, and
  // the operand traffic from L2 per FLOP drops by a third
  // cta2_mink (default 512): activation x weight GEMMs with a long reduction (the qkv dgrad, K = 768; any K = 1024 shape
  // without a fused LayerNorm) also run as CTA pairs: with K >= 512 a tile holds 8+ k-blocks, the accumulator hand-off is
  // amortised, and the per-SM operand ingest drops from 48 to 32 KB per k-block — these shapes were waiting for operands
  // 39-54 % of the MMA warp's time (profiles/r2_gemm_wait_accounting.txt): dgrad_qkv 108-117 -> 99-100 us per launch
  const bool cta2 = g_dbg.cta2 != 0 ||
                    (g_dbg.cta2_wgrad != 0 && a_mn && b_mn && ksplit > 1 && M % (2 * BM) == 0 && N % BN == 0) ||
                    (g_dbg.cta2_mink > 0 && !a_mn && K >= g_dbg.cta2_mink && ksplit <= 1 && !epi->ln_out &&
                     M % (2 * BM) == 0 && N % BN == 0);
  // activation x weight GEMMs (forward, dgrad): clusters of two independent CTAs that share the weight tile of every
  // stage by multicast (CTA2 = 2 in the kernel).  Needs whole 256-row pairs and whole 256-column weight tiles.
  // Opt-in (sc_debug_set("mcb", 1)): parity-green, but no shape got faster -- the weight re-loads are L2 hits that
  // these kernels do not wait for (profiles/README.md).
  const bool mcb = !cta2 && g_dbg.mcb != 0 && !a_mn && ksplit <= 1 && !epi->colsum && M % (2 * BM) == 0 && N % BN == 0 &&
                   M >= 2 * BM * 2;
  p.tiles_m = (cta2 || mcb) ? static_cast<int>((M + 2 * BM - 1) / (2 * BM)) : static_cast<int>((M + BM - 1) / BM);
  p.tiles_n = static_cast<int>((N + BN - 1) / BN);
  p.kb_total = static_cast<int>((K + BK - 1) / BK);
  int ks = ksplit < 1 ? 1 : ksplit;
  if (ks > p.kb_total) ks = p.kb_total;
  p.kb_per_split = (p.kb_total + ks - 1) / ks;
  p.ksplit = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;   // no empty split
  if (p.ksplit > 1) {
    SC_REQUIRE(epi->atomic && epi->out_f32 && !epi->out_bf16 && !epi->bias && !epi->rowadd && !epi->rowgrp &&
                   !epi->resid && !epi->aux_in && !epi->aux_out && epi->act == SC_ACT_NONE,
               SC_E_SHAPE, "sc_gemm_bf16: ksplit > 1 needs a pure atomic fp32 epilogue");
  }
  p.bias = epi->bias;
  p.rowadd = epi->rowadd;
  p.rowadd_period = static_cast<uint32_t>(epi->rowadd_period);
  p.rowgrp = epi->rowgrp;
  p.rowgrp_size = static_cast<uint32_t>(epi->rowgrp_size);
  p.rowgrp_scale = epi->rowgrp_scale;
  p.alpha = epi->alpha;
  p.resid = static_cast<const bf16*>(epi->resid);
  p.aux_in = static_cast<const bf16*>(epi->aux_in);
  p.aux_out = static_cast<bf16*>(epi->aux_out);
  p.out_bf16 = static_cast<bf16*>(epi->out_bf16);
  p.out_f32 = epi->out_f32;
  p.ldo = epi->ldo;
  p.act = epi->act;
  p.aux_in_erf = epi->aux_in_erf;
  p.atomic = epi->atomic;
  p.colsum = epi->colsum;
  p.dbg = static_cast<int>(g_dbg.gemm_dbg);
  p.reverse = sc_take_direction();
  SC_REQUIRE(!epi->colsum || (a_mn == 1 && epi->out_f32 && !epi->out_bf16), SC_E_SHAPE,
             "sc_gemm_bf16: colsum needs a_mn = 1 and an fp32-only output (the wgrad form)");
  p.vec_ok = (N % 8 == 0) && (epi->ldo % 8 == 0) && sc_aligned16(epi->bias) && sc_aligned16(epi->rowadd) &&
             sc_aligned16(epi->rowgrp) && sc_aligned16(epi->resid) && sc_aligned16(epi->aux_in) &&
             sc_aligned16(epi->aux_out) && sc_aligned16(epi->out_bf16) && sc_aligned16(epi->out_f32);

  // UMMA shared-memory descriptor fields
  p.a_lbo = a_mn ? CHUNK_BYTES : 16;   p.a_sbo = 1024;   p.a_kstep = a_mn ? 2048 : 32;
  p.b_lbo = b_mn ? CHUNK_BYTES : 16;   p.b_sbo = 1024;   p.b_kstep = b_mn ? 2048 : 32;
  if (g_dbg.a_lbo) p.a_lbo = (uint32_t)g_dbg.a_lbo;
  if (g_dbg.a_sbo) p.a_sbo = (uint32_t)g_dbg.a_sbo;
  if (g_dbg.a_kstep) p.a_kstep = (uint32_t)g_dbg.a_kstep;
  if (g_dbg.b_lbo) p.b_lbo = (uint32_t)g_dbg.b_lbo;
  if (g_dbg.b_sbo) p.b_sbo = (uint32_t)g_dbg.b_sbo;
  if (g_dbg.b_kstep) p.b_kstep = (uint32_t)g_dbg.b_kstep;

  // skinny operands: load and multiply only the live part of the tile
  p.n_mma = BN; p.n_chunks = BN / 64; p.m_chunks = BM / 64;
  if (!cta2 && p.tiles_n == 1 && N < BN) {
    p.n_mma = static_cast<int>((N + 15) / 16 * 16);
    p.n_chunks = (p.n_mma + 63) / 64;
  }
  if (!cta2 && p.tiles_m == 1 && M <= 64) p.m_chunks = 1;
  p.a_tx = a_mn ? static_cast<uint32_t>(p.m_chunks) * CHUNK_BYTES : A_BYTES;
  p.b_tx = b_mn ? static_cast<uint32_t>(p.n_chunks) * CHUNK_BYTES : static_cast<uint32_t>(p.n_mma) * BK * 2;

  CUtensorMap tmA, tmB;
  if (a_mn == 0) rc = sc_encode_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, BK, BM);
  else           rc = sc_encode_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64, BK);
  if (rc != SC_OK) return rc;
  if (b_mn == 0) rc = sc_encode_tmap_2d_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, BK, (cta2 || mcb) ? BN / 2 : p.n_mma);
  else           rc = sc_encode_tmap_2d_bf16(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64, BK);
  if (rc != SC_OK) return rc;

  // staged (TMA) epilogue whenever the output is a plain aligned bf16 matrix with at most one bf16 input tile
  CUtensorMap tmOut, tmAux, tmIn;
  memset(&tmOut, 0, sizeof(tmOut)); memset(&tmAux, 0, sizeof(tmAux)); memset(&tmIn, 0, sizeof(tmIn));
  p.epi_mode = EPI_GENERIC;
  const bool pos16 = epi->rowadd_bf16 != nullptr && epi->rowadd_period % 32 == 0;
  const int n_in = (epi->resid != nullptr) + (epi->aux_in != nullptr) + (pos16 ? 1 : 0);
  const bool one_in = n_in == 1;
  const bool no_in = n_in == 0;
  SC_REQUIRE(!(epi->rowadd && epi->rowadd_bf16), SC_E_SHAPE, "sc_gemm_bf16: pass rowadd or rowadd_bf16, not both");
  SC_REQUIRE(!epi->rowadd_bf16 || epi->rowadd_period > 0, SC_E_SHAPE, "sc_gemm_bf16: rowadd_period must be > 0");
  bool staged_ok = false;
  // the exact (erf) GELU exists only in the direct-store epilogue: it is an off-by-default spec variant, and its
  // polynomial would cost the staged kinds (one kernel, one register allocation) registers they do not have
  const bool erf_epi = epi->act == SC_ACT_GELU_ERF || (epi->aux_in && epi->aux_in_erf);
  if (!g_dbg.generic_epi && !erf_epi && p.vec_ok && sc_aligned16(epi->rowadd_bf16) && epi->out_bf16 && !epi->out_f32 && p.ksplit == 1 && (no_in || one_in) &&
      !(one_in && epi->aux_out) && N % BN == 0) {
    staged_ok = true;
    p.epi_mode = one_in ? EPI_IN : (epi->aux_out ? EPI_OUT2 : EPI_OUT1);
    p.epi_kind = KIND_GEN;
    if (epi->alpha == 1.0f && !epi->rowadd && !epi->rowgrp) {
      const bool b = epi->bias != nullptr;
      if (!b && epi->act == SC_ACT_NONE && !epi->aux_out && no_in) p.epi_kind = KIND_PLAIN;
      else if (b && epi->act == SC_ACT_NONE && !epi->aux_out && no_in) p.epi_kind = KIND_BIAS;
      else if (b && epi->act == SC_ACT_GELU && epi->aux_out && no_in) p.epi_kind = KIND_FC1;
      else if (b && epi->act == SC_ACT_NONE && (epi->resid || pos16)) p.epi_kind = epi->ln_out ? KIND_RESID_LN : KIND_RESID;
      else if (!b && epi->act == SC_ACT_NONE && epi->aux_in && !epi->aux_in_erf) p.epi_kind = KIND_DGELU;
    }
    const uint64_t ldb_bytes = (uint64_t)epi->ldo * 2;
    if ((rc = sc_encode_tmap_2d_bf16(&tmOut, epi->out_bf16, (uint64_t)N, (uint64_t)M, ldb_bytes, 64, 32)) != SC_OK) return rc;
    if (epi->aux_out &&
        (rc = sc_encode_tmap_2d_bf16(&tmAux, epi->aux_out, (uint64_t)N, (uint64_t)M, ldb_bytes, 64, 32)) != SC_OK) return rc;
    if (epi->ln_out &&
        (rc = sc_encode_tmap_2d_bf16(&tmAux, epi->ln_out, (uint64_t)N, (uint64_t)M, ldb_bytes, 64, 32)) != SC_OK) return rc;
    if (one_in && pos16) {
      p.in_row_period = (int)epi->rowadd_period;
      if ((rc = sc_encode_tmap_2d_bf16(&tmIn, epi->rowadd_bf16, (uint64_t)N, (uint64_t)epi->rowadd_period,
                                       (uint64_t)N * 2, 64, 32)) != SC_OK) return rc;
    } else if (one_in) {
      const void* in = epi->resid ? epi->resid : epi->aux_in;
      if ((rc = sc_encode_tmap_2d_bf16(&tmIn, in, (uint64_t)N, (uint64_t)M, ldb_bytes, 64, 32)) != SC_OK) return rc;
    }
  }
  if (epi->ln_out) {
    SC_REQUIRE(staged_ok && p.epi_kind == KIND_RESID_LN && N == BN && !cta2 && epi->ln_gamma && epi->ln_beta &&
                   sc_aligned16(epi->ln_out),
               SC_E_SHAPE,
               "sc_gemm_bf16: the fused LayerNorm needs N == 256, a bf16 output with bias and one bf16 input tile "
               "(resid or rowadd_bf16), alpha == 1, no activation, ln_gamma and ln_beta");
    p.ln_gamma = epi->ln_gamma; p.ln_beta = epi->ln_beta; p.ln_mean = epi->ln_mean; p.ln_rstd = epi->ln_rstd;
    p.ln_eps = epi->ln_eps;
  }
  SC_REQUIRE(staged_ok || !epi->rowadd_bf16, SC_E_SHAPE,
             "sc_gemm_bf16: rowadd_bf16 needs the staged epilogue (bf16 output, N %% 256 == 0, period %% 32 == 0, "
             "no other bf16 input tile)");

  const long long items = 1LL * p.tiles_m * p.tiles_n * p.ksplit;
  int grid = sc_num_sms();
  if (g_dbg.max_ctas > 0 && grid > g_dbg.max_ctas) grid = (int)g_dbg.max_ctas;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // K <= 256: the tile's MMAs are short and the step is epilogue-bound.  Math-heavy / wide epilogues get two
  // epilogue warpgroups (two warps per scheduler hide each other's dependency stalls, 3 pipeline stages); the light
  // residual epilogue does better with one warpgroup and four staging slots per warp (the next tile's input is
  // prefetched a whole tile ahead, 3 stages).  Longer K: one warpgroup, two slots, 4 stages.
  // (measured on B200, M = 262144: profiles/r1_gemm_epilogue_configs.json)
  const bool short_k = K <= 256 && p.epi_mode != EPI_GENERIC;
  const bool light_in = short_k && p.epi_mode == EPI_IN && p.epi_kind == KIND_RESID;
  const bool fused_ln = p.epi_kind == KIND_RESID_LN && p.epi_mode == EPI_IN;       // needs 1 warpgroup x 4 slots
  // fused LayerNorm: the row split between two warpgroups (2 slots each) unless sc_debug_set("ln_split", 0) asks for the
  // one-warpgroup / 4-slot form
  // (measured: K <= 256 shapes gain 4-7 %, the K = 1024 fc2 shape loses 2-4 % -> short K only; ln_split = 2 forces it)
  p.ln_split = (fused_ln && (g_dbg.ln_split == 2 || (g_dbg.ln_split == 1 && K <= 256))) ? 1 : 0;
  p.ln_prefetch = g_dbg.ln_prefetch != 0;
  const int nwg = fused_ln ? (p.ln_split ? 2 : 1) : (g_dbg.nwg ? (g_dbg.nwg == 2 ? 2 : 1) : ((short_k && !light_in) ? 2 : 1));
  int slots = fused_ln ? (p.ln_split ? 2 : 4) : (g_dbg.slots ? (g_dbg.slots == 4 ? 4 : 2) : (light_in ? 4 : 2));
  // one_slot: output-only staged epilogues of the short-K shapes (qkv, the proj dgrad, ...) with two warpgroups give one of
  // their two staging slots per warp to the pipeline: 4 x 48 KB stages hold a whole K = 256 tile (wait accounting: the MMA
  // warp of the qkv GEMM waited for operands 38 % of its time with both epilogue warpgroups idle half of theirs):
  // qkv 126.0 -> 123.6 us, the K = 256 plain dgrads 56.4 -> 53.1 us per 262144 rows; sc_debug_set("one_slot", 0) = off
  if (!fused_ln && nwg == 2 && slots == 2 && p.epi_mode == EPI_OUT1 && !p.ln_split && g_dbg.one_slot != 0) slots = 1;
  // output-only staged epilogues with two warpgroups: column split as an option (sc_debug_set("col_split", 1 | 2)); off:
  // on the whole step it is neutral (gemm_fwd 6.708 vs 6.705 ms, dgrad 1.96 vs 2.00 ms)
  // (measured: the one-output shapes gain 2-5 %; the two-output fc1 epilogue waits for its own stores at every chunk and
  //  loses 16 % when a warpgroup's two chunks follow each other directly -> EPI_OUT1 only, col_split = 2 forces both)
  if (!fused_ln && !cta2 && nwg == 2 && slots == 2 &&
      ((g_dbg.col_split != 0 && p.epi_mode == EPI_OUT1) || (g_dbg.col_split == 2 && p.epi_mode == EPI_OUT2)))
    p.ln_split = 1;
#define SC_LAUNCH1(AM, BMJ, C2, W, S) launch<AM, BMJ, C2, W, S>(tmA, tmB, tmOut, tmAux, tmIn, p, grid, st)
#define SC_LAUNCH(AM, BMJ)                                                                          \
  (nwg == 2 ? (slots == 4 ? SC_LAUNCH1(AM, BMJ, 0, 2, 4) : (slots == 1 ? SC_LAUNCH1(AM, BMJ, 0, 2, 1) : SC_LAUNCH1(AM, BMJ, 0, 2, 2))) \
            : (slots == 4 ? SC_LAUNCH1(AM, BMJ, 0, 1, 4) : SC_LAUNCH1(AM, BMJ, 0, 1, 2)))
#define SC_LAUNCH_PAIR(AM, BMJ) (nwg == 2 ? SC_LAUNCH1(AM, BMJ, 1, 2, 2) : SC_LAUNCH1(AM, BMJ, 1, 1, 2))
#define SC_LAUNCH_MCB(AM, BMJ)                                                                      \
  (nwg == 2 ? (slots == 4 ? SC_LAUNCH1(AM, BMJ, 2, 2, 4) : SC_LAUNCH1(AM, BMJ, 2, 2, 2))           \
            : (slots == 4 ? SC_LAUNCH1(AM, BMJ, 2, 1, 4) : SC_LAUNCH1(AM, BMJ, 2, 1, 2)))
  if (mcb) {
    int pairs = grid / 2;
    if (items < pairs) pairs = (int)items;
    grid = 2 * pairs;
    return b_mn ? SC_LAUNCH_MCB(0, 1) : SC_LAUNCH_MCB(0, 0);
  }
  if (cta2) {
    int pairs = grid / 2;
    if (items < pairs) pairs = (int)items;
    grid = 2 * pairs;
    if (a_mn == 0 && b_mn == 0) return SC_LAUNCH_PAIR(0, 0);
    if (a_mn == 0 && b_mn == 1) return SC_LAUNCH_PAIR(0, 1);
    if (a_mn == 1 && b_mn == 1) return SC_LAUNCH_PAIR(1, 1);
    return SC_LAUNCH_PAIR(1, 0);
  }
  if (items < grid) grid = (int)items;
  if (a_mn == 0 && b_mn == 0) return SC_LAUNCH(0, 0);
  if (a_mn == 0 && b_mn == 1) return SC_LAUNCH(0, 1);
  if (a_mn == 1 && b_mn == 1) return SC_LAUNCH(1, 1);
  return SC_LAUNCH(1, 0);
#undef SC_LAUNCH_PAIR
#undef SC_LAUNCH_MCB
#undef SC_LAUNCH1
#undef SC_LAUNCH
}
