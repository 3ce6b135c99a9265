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
// CTA2 = 0: one CTA per tile (cta_group::1), B tile 256 x 64 per stage, 3 stages of 48 KB.
// CTA2 = 1: a CTA pair (cluster of 2, cta_group::2) computes a 256 x 256 tile; each CTA stages its own 128 A rows
//           and HALF of the B tile (128 x 64) -> 32 KB stages, 4 of them, and half the B shared-memory traffic per SM.
// NWG = epilogue warpgroups (1 or 2): the second warpgroup costs 32 KB of staging slots = one pipeline stage.
template <int CTA2, int NWG> struct Geo {
  static constexpr int STAGES = (CTA2 ? 4 : 3) + (NWG == 1 ? 1 : 0);
  static constexpr int B_ROWS = CTA2 ? BN / 2 : BN;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM = STAGES * STAGE_BYTES;
  static constexpr int RING = NWG * 4 * 2 * (32 * 128);                   // NWG x 4 warps x 2 slots x 4 KB
  static constexpr int THREADS = 128 + NWG * 128;
};
constexpr int CHUNK_BYTES = 64 * BK * 2;             // 8192: one 64-wide MN-major chunk
constexpr int EPI_SLOT_BYTES = 32 * 128;              // 32 rows x 64 bf16 columns, SWIZZLE_128B
constexpr int EPI_WARPS = 8;                         // two epilogue warpgroups, one per accumulator stage
constexpr int EPI_SLOTS = 2;                         // slots per epilogue warp
constexpr int EPI_RING_BYTES = EPI_WARPS * EPI_SLOTS * EPI_SLOT_BYTES;   // 64 KB
template <int CTA2, int NWG> constexpr int smem_bytes() { return Geo<CTA2, NWG>::SMEM + Geo<CTA2, NWG>::RING + 1024; }
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
constexpr int NUM_THREADS = 128 + EPI_WARPS * 32;   // 384
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
  int atomic;
  int vec_ok;
  int epi_mode;
  int epi_kind;
  float* colsum;                // a_mn = 1 only: colsum[m] += sum_k A(m,k) via an extra MMA against an all-ones tile
  int in_row_period;            // > 0: the TMA input tile is a [period, N] table indexed by row % period
  // UMMA descriptor fields (bytes); defaults set by the host, overridable for bring-up
  uint32_t a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep;
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

template <bool kFast>
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == SC_ACT_GELU) return gelu_fwd<kFast>(x);
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
    x[0] *= gelu_bwd<true>(a0.x); x[1] *= gelu_bwd<true>(a0.y);
    x[2] *= gelu_bwd<true>(a1.x); x[3] *= gelu_bwd<true>(a1.y);
    x[4] *= gelu_bwd<true>(a2.x); x[5] *= gelu_bwd<true>(a2.y);
    x[6] *= gelu_bwd<true>(a3.x); x[7] *= gelu_bwd<true>(a3.y);
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
  if (p.aux_in) x *= gelu_bwd<true>(__bfloat162float(p.aux_in[off]));
  if (p.resid) x += __bfloat162float(p.resid[off]);
  if (p.out_bf16) p.out_bf16[off] = __float2bfloat16(x);
  if (p.out_f32) {
    if (p.atomic) atomicAdd(p.out_f32 + off, x);
    else p.out_f32[off] = x;
  }
}

// 8 columns of one row for the staged epilogue: math only, results packed to bf16.
// KIND != KIND_GEN turns every feature test into a compile-time constant.
template <int KIND>
__device__ __forceinline__ void epilogue_math8(const GemmParams& p, const uint32_t* acc, int col, uint32_t row_mod,
                                               uint32_t row_grp, const uint4& in, uint32_t (&outp)[4],
                                               uint32_t (&auxp)[4]) {
  constexpr bool G = (KIND == KIND_GEN);
  const bool has_bias = G ? (p.bias != nullptr) : (KIND == KIND_BIAS || KIND == KIND_FC1 || KIND == KIND_RESID);
  const bool want_aux = G ? (p.aux_out != nullptr) : (KIND == KIND_FC1);
  const int act = G ? p.act : (KIND == KIND_FC1 ? SC_ACT_GELU : SC_ACT_NONE);
  const bool in_dgelu = G ? (p.aux_in != nullptr) : (KIND == KIND_DGELU);
  const bool in_resid = G ? (p.epi_mode == EPI_IN && p.aux_in == nullptr) : (KIND == KIND_RESID);
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(acc[j]);
  if (G) {
    if (p.alpha != 1.0f) {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] *= p.alpha;
    }
  }
  if (has_bias) {
    float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
    float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
    x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
    x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
  }
  if (G) {
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
  }
  if (want_aux) {
    auxp[0] = pack_bf16(x[0], x[1]); auxp[1] = pack_bf16(x[2], x[3]);
    auxp[2] = pack_bf16(x[4], x[5]); auxp[3] = pack_bf16(x[6], x[7]);
  }
  if (act == SC_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = gelu_fwd<true>(x[j]);
  } else if (G && act == SC_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = tanh_fast(x[j]);
  }
  if (in_dgelu || in_resid) {
    const float2 a0 = unpack_bf16(in.x), a1 = unpack_bf16(in.y), a2 = unpack_bf16(in.z), a3 = unpack_bf16(in.w);
    if (in_dgelu) {
      x[0] *= gelu_bwd<true>(a0.x); x[1] *= gelu_bwd<true>(a0.y); x[2] *= gelu_bwd<true>(a1.x);
      x[3] *= gelu_bwd<true>(a1.y); x[4] *= gelu_bwd<true>(a2.x); x[5] *= gelu_bwd<true>(a2.y);
      x[6] *= gelu_bwd<true>(a3.x); x[7] *= gelu_bwd<true>(a3.y);
    } else {
      x[0] += a0.x; x[1] += a0.y; x[2] += a1.x; x[3] += a1.y;
      x[4] += a2.x; x[5] += a2.y; x[6] += a3.x; x[7] += a3.y;
    }
  }
  outp[0] = pack_bf16(x[0], x[1]); outp[1] = pack_bf16(x[2], x[3]);
  outp[2] = pack_bf16(x[4], x[5]); outp[3] = pack_bf16(x[6], x[7]);
}

// One 128x256 accumulator tile of the staged epilogue for one warp (32 rows): four 64-column chunks, each
// drained in 16-column steps (tcgen05.ld.x16 -> math -> st.shared into a SWIZZLE_128B slot) and then written
// with one (two for EPI_OUT2) TMA store.  Two slots per warp: chunk c uses slot c & 1.  In EPI_IN mode the slot
// first receives the bf16 input tile by TMA load (chunks 0,1 prefetched before the tile, chunks 2,3 as soon as
// the store of chunk c-2 has drained the slot) and is overwritten in place with the result.
template <int KIND>
__device__ __forceinline__ void staged_tile(const GemmParams& p, const CUtensorMap* tmOut, const CUtensorMap* tmAux,
                                            const CUtensorMap* tmIn, uint32_t t_addr, uint32_t ring, uint32_t in_bar0,
                                            int lane, int n0, int row0, int in_row0, uint32_t row_mod,
                                            uint32_t row_grp) {
  const bool has_in = (p.epi_mode == EPI_IN);
  const bool two_out = (p.epi_mode == EPI_OUT2);
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    const int col0 = n0 + c * 64;
    if (col0 >= p.N) break;
    const int s = c & 1;
    const uint32_t slot_o = ring + (two_out ? 1 : s) * EPI_SLOT_BYTES;
    const uint32_t slot_a = ring;                                            // pre-activation copy (EPI_OUT2)
    if (has_in) {
      mbar_wait(in_bar0 + s * 8, static_cast<uint32_t>(c >> 1) & 1u);        // slot s: use 2t + (c >> 1) of this barrier
    } else if (lane == 0) {
      // the TMA store that last read the slot(s) must have drained them
      if (two_out) tma_store_wait_read<0>(); else tma_store_wait_read<1>();
    }
    __syncwarp();
    const uint32_t row_o = slot_o + lane * 128, row_a = slot_a + lane * 128;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
#pragma unroll 1
    for (int h = 0; h < 4; ++h) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(t_addr + c * 64 + h * 16, v);
      const uint32_t pa = ((2u * h) ^ sw) << 4, pb = ((2u * h + 1u) ^ sw) << 4;
      uint4 i0 = make_uint4(0u, 0u, 0u, 0u), i1 = i0;
      if (has_in) {
        i0 = ld_shared_v4(row_o + pa);
        i1 = ld_shared_v4(row_o + pb);
      }
      tmem_ld_wait();
      uint32_t o[4], a[4];
      epilogue_math8<KIND>(p, &v[0], col0 + h * 16, row_mod, row_grp, i0, o, a);
      st_shared_v4(row_o + pa, o[0], o[1], o[2], o[3]);
      if (two_out) st_shared_v4(row_a + pa, a[0], a[1], a[2], a[3]);
      epilogue_math8<KIND>(p, &v[8], col0 + h * 16 + 8, row_mod, row_grp, i1, o, a);
      st_shared_v4(row_o + pb, o[0], o[1], o[2], o[3]);
      if (two_out) st_shared_v4(row_a + pb, a[0], a[1], a[2], a[3]);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (two_out) tma_store_2d(tmAux, slot_a, col0, row0);
      tma_store_2d(tmOut, slot_o, col0, row0);
      tma_store_commit();
      if (has_in && c < 2) {                                                 // refill the slot with chunk c + 2's input
        tma_store_wait_read<0>();
        mbar_arrive_expect_tx(in_bar0 + s * 8, EPI_SLOT_BYTES);
        tma_load_2d(slot_o, tmIn, in_bar0 + s * 8, col0 + 128, in_row0);
      }
    }
  }
}

template <int A_MN, int B_MN, int CTA2, int NWG>
__global__ void __launch_bounds__(128 + NWG * 128, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAux,
                    const __grid_constant__ CUtensorMap tmIn, const GemmParams p) {
  using G = Geo<CTA2, NWG>;
  constexpr int STAGES = G::STAGES;
  constexpr int STAGE_BYTES = G::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ __align__(8) uint64_t epi_in_bar[EPI_WARPS][EPI_SLOTS];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B atoms: 1 KB aligned
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;       // 0 = MMA leader of the pair
  const int ncta = CTA2 ? 2 : 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tmem_full_bar[s]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[s]), 128 * ncta);   // pair: both CTAs' epilogue warpgroups arrive on the leader's
    }
    for (int w = 0; w < EPI_WARPS; ++w)
      for (int s = 0; s < EPI_SLOTS; ++s) mbar_init(smem_u32(&epi_in_bar[w][s]), 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 2) {
    if (CTA2) tmem_alloc_2sm<TMEM_COLS>(smem_u32(&tmem_base_slot));
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
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();            // peer barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // CTA2: p.tiles_m counts 256-row tile pairs; this CTA owns 128-row block 2 * pair + rank of each
  const int num_items = p.tiles_m * p.tiles_n * p.ksplit;
  const int cta_stride = CTA2 ? static_cast<int>(gridDim.x) / 2 : static_cast<int>(gridDim.x);
  const int cta_first = CTA2 ? static_cast<int>(blockIdx.x) / 2 : static_cast<int>(blockIdx.x);
  auto get_item = [&](int item) {
    Item it = decode_item(p, item);
    if (CTA2) it.m_blk = it.m_blk * 2 + static_cast<int>(rank);
    return it;
  };
  const bool single_acc = (p.colsum != nullptr);       // stage 1's columns hold the row-sum accumulator

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cta_first; item < num_items; item += cta_stride) {
        const Item it = get_item(item);
        const int m0 = it.m_blk * BM;
        const int n0 = it.n_blk * BN + static_cast<int>(rank) * G::B_ROWS;   // pair: this CTA's half of B
        for (int kb = it.kb_begin; kb < it.kb_end; ++kb) {
          const uint32_t fb_local = smem_u32(&full_bar[stage]);
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
          const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
          const uint32_t b_dst = a_dst + A_BYTES;
          const int k0 = kb * BK;
          if (CTA2) {
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
            mbar_arrive_expect_tx(fb, STAGE_BYTES);
            if (A_MN == 0) {
              tma_load_2d(a_dst, &tmA, fb, k0, m0);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c) tma_load_2d(a_dst + c * CHUNK_BYTES, &tmA, fb, m0 + 64 * c, k0);
            }
            if (B_MN == 0) {
              tma_load_2d(b_dst, &tmB, fb, k0, n0);
            } else {
#pragma unroll
              for (int c = 0; c < BN / 64; ++c) tma_load_2d(b_dst + c * CHUNK_BYTES, &tmB, fb, n0 + 64 * c, k0);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && rank == 0) {                              // pair: only the leader CTA issues MMAs
      constexpr uint32_t idesc = umma_idesc_bf16(CTA2 ? 2 * BM : BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int item = cta_first; item < num_items; item += cta_stride) {
        const Item it = get_item(item);
        mbar_wait(smem_u32(&tmem_empty_bar[as]), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = it.kb_begin; kb < it.kb_end; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * STAGE_BYTES;
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = umma_smem_desc(a_addr + k * p.a_kstep, p.a_lbo, p.a_sbo);
            const uint64_t bd = umma_smem_desc(b_addr + k * p.b_kstep, p.b_lbo, p.b_sbo);
            const uint32_t acc = (kb > it.kb_begin || k > 0) ? 1u : 0u;
            if (CTA2) umma_bf16_ss_2sm(d_tmem, ad, bd, idesc, acc);
            else umma_bf16_ss(d_tmem, ad, bd, idesc, acc);
            if (single_acc && it.n_blk == 0) {
              constexpr uint32_t idesc_ones = umma_idesc_bf16(CTA2 ? 2 * BM : BM, 16, A_MN, 1);
              const uint64_t od = umma_smem_desc(ones_tile, CHUNK_BYTES, 1024);
              if (CTA2) umma_bf16_ss_2sm(tmem_base + BN, ad, od, idesc_ones, acc);
              else umma_bf16_ss(tmem_base + BN, ad, od, idesc_ones, acc);
            }
          }
          // smem slot reusable (in both CTAs of a pair) when the MMAs retire
          if (CTA2) umma_commit_2sm(smem_u32(&empty_bar[stage])); else umma_commit(smem_u32(&empty_bar[stage]));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        // accumulator complete (published to both CTAs of a pair)
        if (CTA2) umma_commit_2sm(smem_u32(&tmem_full_bar[as])); else umma_commit(smem_u32(&tmem_full_bar[as]));
        if (single_acc) {
          aphase ^= 1u;
        } else {
          as ^= 1;
          if (as == 0) aphase ^= 1u;
        }
      }
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
    const int stride2 = NWG * cta_stride;
    const int first = cta_first + wg * cta_stride;
    // pair: the leader's MMA thread waits for BOTH CTAs' epilogues -> arrive on the leader's barrier
    auto release_acc = [&](int as) {
      const uint32_t local = smem_u32(&tmem_empty_bar[as]);
      if (CTA2) mbar_arrive_cluster(mapa_shared(local, 0)); else mbar_arrive(local);
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
      const uint32_t ring = smem_base + G::SMEM + ew * (EPI_SLOTS * EPI_SLOT_BYTES);
      static_assert(G::RING == NWG * 4 * EPI_SLOTS * EPI_SLOT_BYTES, "ring size");
      const bool has_in = (p.epi_mode == EPI_IN);
      const uint32_t ib0 = smem_u32(&epi_in_bar[ew][0]);
      auto in_row = [&](int row0) { return p.in_row_period > 0 ? row0 % p.in_row_period : row0; };
      if (has_in && lane == 0 && first < num_items) {
        const Item it = get_item(first);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          mbar_arrive_expect_tx(ib0 + c * 8, EPI_SLOT_BYTES);
          tma_load_2d(ring + c * EPI_SLOT_BYTES, &tmIn, ib0 + c * 8, it.n_blk * BN + c * 64, in_row(it.m_blk * BM + q * 32));
        }
      }
      int n = wg;
      for (int item = first; item < num_items; item += stride2, n += NWG) {
        const Item it = get_item(item);
        const int as = n & 1;
        const uint32_t t_addr = t_lane + static_cast<uint32_t>(as * BN);
        const int row0 = it.m_blk * BM + q * 32;
        const int64_t row = min(static_cast<int64_t>(row0) + lane, p.M - 1);   // tail rows: TMA clips the store
        const int n0 = it.n_blk * BN;
        uint32_t row_mod = 0, row_grp = 0;
        if (p.rowadd) row_mod = static_cast<uint32_t>(row) % p.rowadd_period;
        if (p.rowgrp) row_grp = static_cast<uint32_t>(row) / p.rowgrp_size;
        mbar_wait(smem_u32(&tmem_full_bar[as]), static_cast<uint32_t>((n >> 1) & 1));
        __syncwarp();
        tc_fence_after();
        const int ir0 = in_row(row0);
        switch (p.epi_kind) {
          case KIND_PLAIN: staged_tile<KIND_PLAIN>(p, &tmOut, &tmAux, &tmIn, t_addr, ring, ib0, lane, n0, row0, ir0, row_mod, row_grp); break;
          case KIND_BIAS:  staged_tile<KIND_BIAS>(p, &tmOut, &tmAux, &tmIn, t_addr, ring, ib0, lane, n0, row0, ir0, row_mod, row_grp); break;
          case KIND_FC1:   staged_tile<KIND_FC1>(p, &tmOut, &tmAux, &tmIn, t_addr, ring, ib0, lane, n0, row0, ir0, row_mod, row_grp); break;
          case KIND_RESID: staged_tile<KIND_RESID>(p, &tmOut, &tmAux, &tmIn, t_addr, ring, ib0, lane, n0, row0, ir0, row_mod, row_grp); break;
          case KIND_DGELU: staged_tile<KIND_DGELU>(p, &tmOut, &tmAux, &tmIn, t_addr, ring, ib0, lane, n0, row0, ir0, row_mod, row_grp); break;
          default:         staged_tile<KIND_GEN>(p, &tmOut, &tmAux, &tmIn, t_addr, ring, ib0, lane, n0, row0, ir0, row_mod, row_grp); break;
        }
        tc_fence_before();
        release_acc(as);
        // prefetch the first two input chunks of this warpgroup's next tile once the stores have drained the slots
        const int next = item + stride2;
        if (has_in && lane == 0 && next < num_items) {
          tma_store_wait_read<0>();
          const Item nt = get_item(next);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            mbar_arrive_expect_tx(ib0 + c * 8, EPI_SLOT_BYTES);
            tma_load_2d(ring + c * EPI_SLOT_BYTES, &tmIn, ib0 + c * 8, nt.n_blk * BN + c * 64, in_row(nt.m_blk * BM + q * 32));
          }
        }
      }
      if (lane == 0) tma_store_wait_all();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();            // nobody tears down while the peer may still signal / read it
  if (warp == 2) {
    tc_fence_after();
    if (CTA2) tmem_dealloc_2sm<TMEM_COLS>(tmem_base); else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

struct DebugOverrides {
  int64_t a_lbo = 0, a_sbo = 0, a_kstep = 0, b_lbo = 0, b_sbo = 0, b_kstep = 0, max_ctas = 0, generic_epi = 0, cta2 = 0, nwg = 0;
};
DebugOverrides g_dbg;

template <int A_MN, int B_MN, int CTA2, int NWG>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmAux,
           const CUtensorMap& tmIn, const GemmParams& p, int grid, cudaStream_t st) {
  static bool attr_set = false;
  auto kern = gemm_tcgen05_kernel<A_MN, B_MN, CTA2, NWG>;
  constexpr int smem = smem_bytes<CTA2, NWG>();
  constexpr int threads = Geo<CTA2, NWG>::THREADS;
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
    kern<<<grid, threads, smem, st>>>(tmA, tmB, tmOut, tmAux, tmIn, p);
  }
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_gemm_bf16");
  return SC_OK;
}

}  // namespace

void sc_attn_debug_set_fwd_stages(int v);

extern "C" int sc_debug_set(const char* key, int64_t value) {
  if (!key) return SC_E_SHAPE;
  if (!strcmp(key, "attn_fwd_stages")) { sc_attn_debug_set_fwd_stages((int)value); return SC_OK; }
  if (!strcmp(key, "a_lbo")) g_dbg.a_lbo = value;
  else if (!strcmp(key, "a_sbo")) g_dbg.a_sbo = value;
  else if (!strcmp(key, "a_kstep")) g_dbg.a_kstep = value;
  else if (!strcmp(key, "b_lbo")) g_dbg.b_lbo = value;
  else if (!strcmp(key, "b_sbo")) g_dbg.b_sbo = value;
  else if (!strcmp(key, "b_kstep")) g_dbg.b_kstep = value;
  else if (!strcmp(key, "max_ctas")) g_dbg.max_ctas = value;
  else if (!strcmp(key, "generic_epi")) g_dbg.generic_epi = value;
  else if (!strcmp(key, "cta2")) g_dbg.cta2 = value;
  else if (!strcmp(key, "nwg")) g_dbg.nwg = value;
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
  SC_REQUIRE(epi->act >= SC_ACT_NONE && epi->act <= SC_ACT_TANH, SC_E_SHAPE, "sc_gemm_bf16: bad act");

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  const bool cta2 = g_dbg.cta2 != 0;
  p.tiles_m = cta2 ? static_cast<int>((M + 2 * BM - 1) / (2 * BM)) : static_cast<int>((M + BM - 1) / BM);
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
  p.atomic = epi->atomic;
  p.colsum = epi->colsum;
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

  CUtensorMap tmA, tmB;
  if (a_mn == 0) rc = sc_encode_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, BK, BM);
  else           rc = sc_encode_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64, BK);
  if (rc != SC_OK) return rc;
  if (b_mn == 0) rc = sc_encode_tmap_2d_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, BK, cta2 ? BN / 2 : BN);
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
  if (!g_dbg.generic_epi && p.vec_ok && sc_aligned16(epi->rowadd_bf16) && epi->out_bf16 && !epi->out_f32 && p.ksplit == 1 && (no_in || one_in) &&
      !(one_in && epi->aux_out) && N % BN == 0) {
    staged_ok = true;
    p.epi_mode = one_in ? EPI_IN : (epi->aux_out ? EPI_OUT2 : EPI_OUT1);
    p.epi_kind = KIND_GEN;
    if (epi->alpha == 1.0f && !epi->rowadd && !epi->rowgrp) {
      const bool b = epi->bias != nullptr;
      if (!b && epi->act == SC_ACT_NONE && !epi->aux_out && no_in) p.epi_kind = KIND_PLAIN;
      else if (b && epi->act == SC_ACT_NONE && !epi->aux_out && no_in) p.epi_kind = KIND_BIAS;
      else if (b && epi->act == SC_ACT_GELU && epi->aux_out && no_in) p.epi_kind = KIND_FC1;
      else if (b && epi->act == SC_ACT_NONE && (epi->resid || pos16)) p.epi_kind = KIND_RESID;
      else if (!b && epi->act == SC_ACT_NONE && epi->aux_in) p.epi_kind = KIND_DGELU;
    }
    const uint64_t ldb_bytes = (uint64_t)epi->ldo * 2;
    if ((rc = sc_encode_tmap_2d_bf16(&tmOut, epi->out_bf16, (uint64_t)N, (uint64_t)M, ldb_bytes, 64, 32)) != SC_OK) return rc;
    if (epi->aux_out &&
        (rc = sc_encode_tmap_2d_bf16(&tmAux, epi->aux_out, (uint64_t)N, (uint64_t)M, ldb_bytes, 64, 32)) != SC_OK) return rc;
    if (one_in && pos16) {
      p.in_row_period = (int)epi->rowadd_period;
      if ((rc = sc_encode_tmap_2d_bf16(&tmIn, epi->rowadd_bf16, (uint64_t)N, (uint64_t)epi->rowadd_period,
                                       (uint64_t)N * 2, 64, 32)) != SC_OK) return rc;
    } else if (one_in) {
      const void* in = epi->resid ? epi->resid : epi->aux_in;
      if ((rc = sc_encode_tmap_2d_bf16(&tmIn, in, (uint64_t)N, (uint64_t)M, ldb_bytes, 64, 32)) != SC_OK) return rc;
    }
  }
  SC_REQUIRE(staged_ok || !epi->rowadd_bf16, SC_E_SHAPE,
             "sc_gemm_bf16: rowadd_bf16 needs the staged epilogue (bf16 output, N %% 256 == 0, period %% 32 == 0, "
             "no other bf16 input tile)");

  const long long items = 1LL * p.tiles_m * p.tiles_n * p.ksplit;
  int grid = sc_num_sms();
  if (g_dbg.max_ctas > 0 && grid > g_dbg.max_ctas) grid = (int)g_dbg.max_ctas;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nwg = g_dbg.nwg == 2 ? 2 : 1;          // default: one epilogue warpgroup + the extra pipeline stage
#define SC_LAUNCH(AM, BMJ, C2)                                                                     \
  (nwg == 2 ? launch<AM, BMJ, C2, 2>(tmA, tmB, tmOut, tmAux, tmIn, p, grid, st)                    \
            : launch<AM, BMJ, C2, 1>(tmA, tmB, tmOut, tmAux, tmIn, p, grid, st))
  if (cta2) {
    int pairs = grid / 2;
    if (items < pairs) pairs = (int)items;
    grid = 2 * pairs;
    if (a_mn == 0 && b_mn == 0) return SC_LAUNCH(0, 0, 1);
    if (a_mn == 0 && b_mn == 1) return SC_LAUNCH(0, 1, 1);
    if (a_mn == 1 && b_mn == 1) return SC_LAUNCH(1, 1, 1);
    return SC_LAUNCH(1, 0, 1);
  }
  if (items < grid) grid = (int)items;
  if (a_mn == 0 && b_mn == 0) return SC_LAUNCH(0, 0, 0);
  if (a_mn == 0 && b_mn == 1) return SC_LAUNCH(0, 1, 0);
  if (a_mn == 1 && b_mn == 1) return SC_LAUNCH(1, 1, 0);
  return SC_LAUNCH(1, 0, 0);
#undef SC_LAUNCH
}
