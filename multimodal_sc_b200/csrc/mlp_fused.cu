// mlp_fused.cu — the transformer MLP as ONE tcgen05 kernel per direction (sm_100a).
//
//   forward  (sc_mlp_fwd):        out  = resid + W2 . gelu(W1 . x + b1) + b2          (SPEC v0 rows a1/a2/a5, block MLP)
//   backward (sc_mlp_bwd_dgrad):  dhpre = (dy . W2) o gelu'(hpre),   dx = dhpre . W1  (the two dgrad GEMMs of the MLP)
//
// Both are the same pipeline: GEMM1 (K = 256) produces a 128 x 128 chunk of the hidden activation in TMEM, the
// epilogue warps apply the element-wise map and write the chunk as bf16 into shared memory in the K-major
// SWIZZLE_128B layout, where it is (a) the A operand of GEMM2, which accumulates the 128 x 256 output over the
// eight hidden chunks, and (b) the source of the TMA store that saves it (h forward, dhpre backward).  The
// 128 x 1024 hidden tile never makes the HBM round trip between the two GEMMs, x / dy are read once per tile.
//
// One CTA per SM, persistent over 128-row tiles, 384 threads:
//   warp 0 / lane 0 : weight producer    — W1 / W2 chunks through a 3 x 32 KB TMA ring (every CTA streams the same
//                                          1 MB of weights per tile; they stay L2-resident)
//   warp 3 / lane 0 : activation producer — the x (dy) tile, backward: the hpre chunks, loaded INTO the hidden buffer
//   warp 1 / lane 0 : MMA issuer          — order G1(0) G1(1) G2(0) G1(2) G2(1) ... so the tensor pipe always has
//                                          the next GEMM1 queued while the epilogue works on the previous chunk
//   warp 2 / lane 0 : TMEM allocator (512 columns: two 128-col GEMM1 accumulators + the 256-col GEMM2 accumulator), then
//                     chunk storer — h / dhpre by TMA straight from the operand buffer, frees the buffer afterwards
//   warps 4-7 / 8-11: epilogue warpgroups — warpgroup w owns GEMM1 accumulator w and hidden buffer w (even / odd
//                                          chunks); warpgroup 1 also drains the output accumulator once per tile
// Shared memory: x tile 64 KB + two hidden buffers 2 x 32 KB + weight ring 96 KB + bias tiles 2 KB.  Nothing is
// staged for the epilogue: hpre and the output leave the registers with 32-byte-per-row vector stores (full
// sectors), h / dhpre leave the operand buffer by TMA, the residual tile is prefetched to L2 by the producer.
#include "common.cuh"

int sc_mlp_bwd_pair_launch(const void* dy, const void* w1, const void* w2, const void* hpre, void* dhpre, void* dx, int64_t M,
                           int R, int dbg, cudaStream_t st);      // mlp_bwd_pair.cu

namespace {

constexpr int BM = 128;                 // rows per tile
constexpr int DM = 256;                 // model width (GEMM1 K, GEMM2 N)
constexpr int HC = 128;                 // hidden columns per chunk (GEMM1 N, GEMM2 K)
constexpr int KB_BYTES = BM * 64 * 2;   // one 128-row x 64-k K-major SWIZZLE_128B block: 16 KB
constexpr int XA_BYTES = 4 * KB_BYTES;  // x / dy tile: 4 k-blocks
constexpr int HB_BYTES = 2 * KB_BYTES;  // one hidden chunk: 2 k-blocks
constexpr int STAGE_BYTES = 32768;      // weight ring stage: GEMM1: 2 k-blocks of 128 x 64; GEMM2: 1 k-block of 256 x 64
constexpr int STAGES = 3;
constexpr int CHUNK8K = 8192;           // one 64-wide MN-major chunk (64 k-rows x 128 B)
constexpr int OFF_XA = 0;
constexpr int OFF_HB = OFF_XA + XA_BYTES;
constexpr int OFF_RING = OFF_HB + 2 * HB_BYTES;
constexpr int OFF_BIAS1 = OFF_RING + STAGES * STAGE_BYTES;      // 2 warpgroups x 128 floats
constexpr int OFF_BIAS2 = OFF_BIAS1 + 2 * HC * 4;               // 256 floats
constexpr int SMEM_TOTAL = OFF_BIAS2 + DM * 4;                  // 231424 = 227 KB - 1 KB of static shared memory
static_assert(SMEM_TOTAL <= 232448 - 1024, "shared memory budget");
constexpr int THREADS = 384;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t ACC2_COL = 256;

struct MlpParams {
  int64_t M;
  int R;                     // hidden width (multiple of 128)
  int tiles;
  const float* b1;
  const float* b2;
  const bf16* resid;         // fwd, optional
  bf16* hpre;                // fwd, optional: pre-activation copy [M, R]
  bf16* out;                 // fwd: out [M, 256];  bwd: dx [M, 256]
  // fwd, optional (inference form, no h / hpre saves): LayerNorm of the output row as a second output — the LayerNorm that
  // consumes the block's output (next block's LN1 / the stack's final norm) rides on this kernel instead of its own launch
  const float* ln_gamma;
  const float* ln_beta;
  bf16* ln_out;
  float* ln_mean;
  float* ln_rstd;
  float ln_eps;
  int store_h;               // fwd: h is saved (TMA store from the operand buffer);  bwd: always 1 (dhpre)
  int reverse;               // 1 = row tiles are walked in descending order
  int dbg;                   // sc_debug_set("mlp_dbg", bits): 1 / 2 = L2 prefetch variants (negative), 4 = no weight multicast
};

// per-role wait accounting (debug builds: SC_MLP_ACCT=1 python -m multimodal_sc_b200.build --force, then
// sc_debug_set("mlp_dbg", 8)): CTA 0 prints, per role, its total cycles and the cycles inside each barrier wait
#ifdef SC_MLP_ACCT
#define MLP_WAIT(bar, par, ctr) do { const long long _t = clock64(); mbar_wait(bar, par); (ctr) += clock64() - _t; } while (0)
constexpr bool kMlpAcct = true;
#else
#define MLP_WAIT(bar, par, ctr) mbar_wait(bar, par)
constexpr bool kMlpAcct = false;
#endif

__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_global_nc_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// ---- packed fp32x2 math (same forms as the GEMM epilogue) ----
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
__device__ __forceinline__ uint64_t unpack2_bf16(uint32_t u) { return pk2u(u << 16, u & 0xFFFF0000u); }
__device__ __forceinline__ uint64_t gelu_fwd2(uint64_t x) {
  const uint64_t k0 = pk2(SC_GELU_C0, SC_GELU_C0), k01 = pk2(SC_GELU_C0 * SC_GELU_C1, SC_GELU_C0 * SC_GELU_C1);
  const uint64_t half = pk2(0.5f, 0.5f);
  const uint64_t u = mul2(x, fma2(mul2(x, x), k01, k0));
  const uint64_t hx = mul2(x, half);
  return fma2(hx, tanh2(u), hx);
}
__device__ __forceinline__ uint64_t neg2(uint64_t a) { return a ^ 0x8000000080000000ull; }   // sign flips: ALU, not FMA pipe
// d/dx gelu_tanh(x) = 0.5 (1 + t) + 0.5 x (1 - t^2) u'(x),  t = tanh(u),  u = k0 x + k01 x^3,  u' = k0 + 3 k01 x^2.
// Nine packed operations per pair (+ tanh): the halves are folded into the constants, 1 - t^2 and 0.5 (1 + t) are one
// FMA each (the chunk epilogue of the fused backward kernel is bound by this arithmetic: ~70 of its ~115 instructions per
// 16-column step).  1 - t^2 instead of (1 + t)(1 - t) loses relative accuracy only where the term is < 1e-5 of the result.
__device__ __forceinline__ uint64_t gelu_bwd2(uint64_t x) {
  const uint64_t k0 = pk2(SC_GELU_C0, SC_GELU_C0), k01 = pk2(SC_GELU_C0 * SC_GELU_C1, SC_GELU_C0 * SC_GELU_C1);
  const uint64_t c0 = pk2(0.5f * SC_GELU_C0, 0.5f * SC_GELU_C0);
  const uint64_t c1 = pk2(1.5f * SC_GELU_C0 * SC_GELU_C1, 1.5f * SC_GELU_C0 * SC_GELU_C1);
  const uint64_t half = pk2(0.5f, 0.5f), one = pk2(1.0f, 1.0f);
  const uint64_t x2 = mul2(x, x);
  const uint64_t t = tanh2(mul2(x, fma2(x2, k01, k0)));
  const uint64_t hxdu = mul2(x, fma2(x2, c1, c0));            // 0.5 x u'(x)
  const uint64_t omt2 = fma2(neg2(t), t, one);                // 1 - t^2
  return fma2(hxdu, omt2, fma2(t, half, half));
}

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  const uint4 w = ld_shared_v4(addr);
  return make_float4(__uint_as_float(w.x), __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w));
}

// MC = 1 (backward only): the kernel runs as clusters of two CTAs that share every weight stage.  Each CTA loads one
// half of the stage and multicasts it into both shared memories; a stage is released when BOTH MMA warps have
// committed (w_empty counts 2, the commit is multicast).  The weights are 1 MB of the 1.6 MB that a 128-row tile pulls
// through L2, and the chip-wide L2 throughput (~6300 B/clk) is what bounds this kernel.
template <int BWD, int MC>
__global__ void __launch_bounds__(THREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWa,
                 const __grid_constant__ CUtensorMap tmWb, const __grid_constant__ CUtensorMap tmH,
                 const __grid_constant__ CUtensorMap tmHin, const __grid_constant__ CUtensorMap tmResid,
                 const MlpParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t x_full, x_empty;
  __shared__ __align__(8) uint64_t w_full[STAGES], w_empty[STAGES];
  __shared__ __align__(8) uint64_t acc1_full[2], acc1_empty[2];
  __shared__ __align__(8) uint64_t h_full[2], h_empty[2], hin_full[2];
  __shared__ __align__(8) uint64_t acc2_full, acc2_empty;
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("semcom: mlp_fused dynamic shared memory base 0x%x is not 1 KB aligned\n", smem_base);
    __trap();
  }
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NC = p.R / HC;                                  // hidden chunks per tile

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&x_full), 1);
    mbar_init(smem_u32(&x_empty), 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&w_full[s]), 1);
      mbar_init(smem_u32(&w_empty[s]), MC ? 2 : 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&acc1_full[b]), 1);
      mbar_init(smem_u32(&acc1_empty[b]), 128);
      mbar_init(smem_u32(&h_full[b]), 128);
      mbar_init(smem_u32(&h_empty[b]), 2);                  // GEMM2 retired (tcgen05.commit) + TMA store has read the buffer
      mbar_init(smem_u32(&hin_full[b]), 1);
    }
    mbar_init(smem_u32(&acc2_full), 1);
    mbar_init(smem_u32(&acc2_empty), 256);                   // both warpgroups drain half of the output tile
    fence_mbar_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWa);
    tma_prefetch_desc(&tmWb);
    tma_prefetch_desc(&tmH);
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(smem_u32(&tmem_base_slot));
  pdl_sync();                                               // no global-memory access above this line (common.cuh)
  if (!BWD) {                                               // b2 is the same for every tile
    for (int i = threadIdx.x; i < DM; i += THREADS)
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem_base + OFF_BIAS2 + i * 4), "f"(p.b2 ? __ldg(p.b2 + i) : 0.f) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();                               // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t crank = MC ? cluster_ctarank() : 0u;

  const uint32_t xa = smem_base + OFF_XA;
  const uint32_t ring = smem_base + OFF_RING;
  auto hb = [&](int b) { return smem_base + OFF_HB + b * HB_BYTES; };

  if (warp == 0) {
    // ===================== weight producer =====================
    // converged warp, one elected lane issues the TMA loads; the next stage's empty barrier is tested ahead
    {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t ready = mbar_test(smem_u32(&w_empty[0]), 1u);
      long long a_w = 0;
      const long long a_t0 = clock64();
      const bool acct = kMlpAcct && (p.dbg & 8) && blockIdx.x == 0;
      auto next_stage = [&]() -> uint32_t {
        if (!ready) MLP_WAIT(smem_u32(&w_empty[stage]), phase ^ 1u, a_w);
        const int nstage = (stage + 1 == STAGES) ? 0 : stage + 1;
        const uint32_t nphase = (stage + 1 == STAGES) ? (phase ^ 1u) : phase;
        ready = mbar_test(smem_u32(&w_empty[nstage]), nphase ^ 1u);
        return ring + stage * STAGE_BYTES;
      };
      auto advance = [&]() { if (++stage == STAGES) { stage = 0; phase ^= 1u; } };
      auto load_g1 = [&](int j) {                           // two stages: k-blocks {0,1}, {2,3} of hidden chunk j
        for (int t = 0; t < 2; ++t) {
          const uint32_t dst = next_stage();
          const uint32_t fb = smem_u32(&w_full[stage]);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(fb, STAGE_BYTES);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int kb = 2 * t + u;
              if (!BWD) {                                   // W1 [R, 256] K-major: 128 hidden rows x 64 k
                tma_load_2d(dst + u * KB_BYTES, &tmWa, fb, kb * 64, j * HC);
              } else if (MC) {                              // this CTA's half of the stage, delivered to both CTAs
                if (u == static_cast<int>(crank)) {
                  tma_load_2d_mc(dst + u * KB_BYTES, &tmWa, fb, j * HC, kb * 64, 3);
                  tma_load_2d_mc(dst + u * KB_BYTES + CHUNK8K, &tmWa, fb, j * HC + 64, kb * 64, 3);
                }
              } else {                                      // W2 [256, R] read MN-major: 64 k-rows x (2 x 64) hidden columns
                tma_load_2d(dst + u * KB_BYTES, &tmWa, fb, j * HC, kb * 64);
                tma_load_2d(dst + u * KB_BYTES + CHUNK8K, &tmWa, fb, j * HC + 64, kb * 64);
              }
            }
          }
          __syncwarp();
          advance();
        }
      };
      auto load_g2 = [&](int j) {                           // two stages: hidden k-blocks {0}, {1} of chunk j, all 256 outputs
        for (int t = 0; t < 2; ++t) {
          const uint32_t dst = next_stage();
          const uint32_t fb = smem_u32(&w_full[stage]);
          const int k0 = j * HC + t * 64;
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(fb, STAGE_BYTES);
            if (!BWD) {                                     // W2 [256, R] K-major: 256 rows x 64 k
              tma_load_2d(dst, &tmWb, fb, k0, 0);
            } else if (MC) {
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                const int cc = 2 * static_cast<int>(crank) + c;
                tma_load_2d_mc(dst + cc * CHUNK8K, &tmWb, fb, 64 * cc, k0, 3);
              }
            } else {                                        // W1 [R, 256] read MN-major: 64 k-rows x (4 x 64) columns
#pragma unroll
              for (int c = 0; c < 4; ++c) tma_load_2d(dst + c * CHUNK8K, &tmWb, fb, 64 * c, k0);
            }
          }
          __syncwarp();
          advance();
        }
      };
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        for (int j = 0; j < NC; ++j) {
          load_g1(j);
          if (j >= 1) load_g2(j - 1);
        }
        load_g2(NC - 1);
      }
      if (acct && lane == 0) printf("mlp acct: weight producer total %lld clk, waiting for a free stage %lld\n", clock64() - a_t0, a_w);
    }
  } else if (warp == 3) {
    // ===================== activation producer =====================
    if (lane == 0) {
      uint32_t t_par = 0;
      uint32_t nuse[2] = {0u, 0u};
      long long a_x = 0, a_h = 0;
      const long long a_t0 = clock64();
      const bool acct = kMlpAcct && (p.dbg & 8) && blockIdx.x == 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, t_par ^= 1u) {
        const int row0 = (p.reverse ? p.tiles - 1 - tile : tile) * BM;
        MLP_WAIT(smem_u32(&x_empty), t_par ^ 1u, a_x);
        mbar_arrive_expect_tx(smem_u32(&x_full), XA_BYTES);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(xa + kb * KB_BYTES, &tmX, smem_u32(&x_full), kb * 64, row0);
        if (!BWD && p.resid) {
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) tma_prefetch_l2_2d(&tmResid, kb * 64, row0);
        }
        // opt-in (mlp_dbg bit 1): the NEXT tile's inputs (x / dy, in backward its eight hpre chunks) requested into L2
        // a whole tile ahead.  Measured negative (3.08 -> 3.26 ms per step, profiles/README.md).
        const int next_tile = tile + static_cast<int>(gridDim.x);
        if ((p.dbg & 1) && next_tile < p.tiles) {
          const int nrow0 = (p.reverse ? p.tiles - 1 - next_tile : next_tile) * BM;
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) tma_prefetch_l2_2d(&tmX, kb * 64, nrow0);
          if (BWD) {
            for (int c = 0; c < 2 * NC; ++c) tma_prefetch_l2_2d(&tmHin, c * 64, nrow0);
          }
        }
        if (BWD) {
          for (int j = 0; j < NC; ++j) {                    // hpre chunk j -> hidden buffer j & 1 (transformed in place)
            const int b = j & 1;
            // mlp_dbg bit 2: the chunk that will be loaded two turns from now is requested into L2 while this one
            // waits for its buffer (short distance: ~2 chunk times, ~100 KB per CTA in flight)
            if (p.dbg & 2) {
              const int jn = j + 2;
              const int ptile = jn < NC ? tile : tile + static_cast<int>(gridDim.x);
              if (ptile < p.tiles) {
                const int prow0 = (p.reverse ? p.tiles - 1 - ptile : ptile) * BM;
                tma_prefetch_l2_2d(&tmHin, (jn % NC) * HC, prow0);
                tma_prefetch_l2_2d(&tmHin, (jn % NC) * HC + 64, prow0);
              }
            }
            MLP_WAIT(smem_u32(&h_empty[b]), (nuse[b] & 1u) ^ 1u, a_h);
            ++nuse[b];
            mbar_arrive_expect_tx(smem_u32(&hin_full[b]), HB_BYTES);
            tma_load_2d(hb(b), &tmHin, smem_u32(&hin_full[b]), j * HC, row0);
            tma_load_2d(hb(b) + KB_BYTES, &tmHin, smem_u32(&hin_full[b]), j * HC + 64, row0);
          }
        }
      }
      if (acct) printf("mlp acct: activation producer total %lld clk, waiting x_empty %lld, h_empty (buffer free) %lld\n", clock64() - a_t0, a_x, a_h);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // All 32 lanes run the loop converged; one elected lane issues (elect_one_sync): inside `if (lane == 0)` ptxas
    // wraps every UTCHMMA in an R2UR/ELECT waterfall of ~48 instructions, more than the 66 cycles a 128x128x16 MMA
    // takes.  The weight stage that is needed NEXT is tested before the current stage's MMAs are issued, so the
    // barrier round trip hides behind them.
    {
      constexpr uint32_t idesc1 = umma_idesc_bf16(BM, HC, 0, BWD ? 1 : 0);
      constexpr uint32_t idesc2 = umma_idesc_bf16(BM, DM, 0, BWD ? 1 : 0);
      constexpr uint32_t b_lbo = BWD ? CHUNK8K : 16;
      constexpr uint32_t b_kstep16 = (BWD ? 2048 : 32) >> 4;
      // descriptors carry the 18-bit CTA-local address only: inside a cluster the shared::cta window of rank r starts
      // at r << 24, and those upper bits would spill into the descriptor's LBO field (seen as wrong second 64-column
      // chunks of MN-major operands on rank 1)
      const uint64_t a_desc0 = umma_smem_desc(0, 16, 1024);
      const uint64_t b_desc0 = umma_smem_desc(0, b_lbo, 1024);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t n1[2] = {0u, 0u}, nh[2] = {0u, 0u};
      uint32_t t_par = 0;
      uint32_t ready = mbar_test(smem_u32(&w_full[0]), 0u);
      int nstage = 0;
      uint32_t nphase = 0;
      long long a_w = 0, a_a1 = 0, a_hf = 0, a_a2 = 0, a_x = 0;
      const long long a_t0 = clock64();
      const bool acct = kMlpAcct && (p.dbg & 8) && blockIdx.x == 0;
      // wait for the current weight stage (it was tested one stage ahead); note which stage comes next
      auto acquire_stage = [&]() -> uint32_t {
        if (!ready) MLP_WAIT(smem_u32(&w_full[stage]), phase, a_w);
        nstage = (stage + 1 == STAGES) ? 0 : stage + 1;
        nphase = (stage + 1 == STAGES) ? (phase ^ 1u) : phase;
        return ring + stage * STAGE_BYTES;
      };
      auto advance = [&]() { stage = nstage; phase = nphase; };
      auto gemm1 = [&](int j) {
        const int b = j & 1;
        MLP_WAIT(smem_u32(&acc1_empty[b]), (n1[b] & 1u) ^ 1u, a_a1);
        ++n1[b];
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(b * HC);
        for (int t = 0; t < 2; ++t) {
          const uint32_t wb = acquire_stage();
          tc_fence_after();
          const uint64_t ad = a_desc0 + (((xa + 2 * t * KB_BYTES) & 0x3FFFFu) >> 4);
          const uint64_t bd = b_desc0 + ((wb & 0x3FFFFu) >> 4);
          uint32_t r = 0;
          if (elect_one_sync()) {
            // eight 128x128x16 MMAs: the next weight stage is tested ahead of the first four, the commit follows the last
            r = umma_kblock4_bf16_ss<true, false>(d_tmem, ad, bd, 2u, b_kstep16, idesc1, t > 0 ? 1u : 0u,
                                                  smem_u32(&w_empty[stage]), smem_u32(&w_full[nstage]), nphase);
            umma_kblock4_bf16_ss<false, true, MC != 0>(d_tmem, ad + (KB_BYTES >> 4), bd + (KB_BYTES >> 4), 2u, b_kstep16,
                                                       idesc1, 1u, smem_u32(&w_empty[stage]), 0u, 0u);
          }
          ready = __shfl_sync(0xffffffffu, r, 0);
          advance();
        }
        if (elect_one_sync()) {
          umma_commit(smem_u32(&acc1_full[b]));
          if (j == NC - 1) umma_commit(smem_u32(&x_empty));             // x tile free once the last GEMM1 retires
        }
        __syncwarp();
      };
      auto gemm2 = [&](int j) {
        const int b = j & 1;
        MLP_WAIT(smem_u32(&h_full[b]), nh[b] & 1u, a_hf);
        ++nh[b];
        if (j == 0) MLP_WAIT(smem_u32(&acc2_empty), t_par ^ 1u, a_a2);      // previous tile's output drained
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ACC2_COL;
        for (int t = 0; t < 2; ++t) {
          const uint32_t wb = acquire_stage();
          tc_fence_after();
          const uint64_t ad = a_desc0 + (((hb(b) + t * KB_BYTES) & 0x3FFFFu) >> 4);
          const uint64_t bd = b_desc0 + ((wb & 0x3FFFFu) >> 4);
          uint32_t r = 0;
          if (elect_one_sync())
            r = umma_kblock4_bf16_ss<true, true, MC != 0>(d_tmem, ad, bd, 2u, b_kstep16, idesc2, (j > 0 || t > 0) ? 1u : 0u,
                                                          smem_u32(&w_empty[stage]), smem_u32(&w_full[nstage]), nphase);
          ready = __shfl_sync(0xffffffffu, r, 0);
          advance();
        }
        if (elect_one_sync()) umma_commit(smem_u32(&h_empty[b]));
        __syncwarp();
      };
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, t_par ^= 1u) {
        MLP_WAIT(smem_u32(&x_full), t_par, a_x);
        tc_fence_after();
        for (int j = 0; j < NC; ++j) {
          gemm1(j);
          if (j >= 1) gemm2(j - 1);
        }
        gemm2(NC - 1);
        if (elect_one_sync()) umma_commit(smem_u32(&acc2_full));
        __syncwarp();
      }
      if (acct && lane == 0)
        printf("mlp acct: mma total %lld clk, waiting weights %lld, acc1_empty %lld, h_full (chunk transformed) %lld, acc2_empty %lld, x_full %lld\n",
               clock64() - a_t0, a_w, a_a1, a_hf, a_a2, a_x);
    }
  } else if (warp == 2) {
    // ===================== chunk storer =====================
    // saves every finished hidden chunk (h forward, dhpre backward) straight from the operand buffer by TMA and frees
    // the buffer once the store has read it.  A thread of its own: the epilogue warpgroups never wait for a store.
    if (lane == 0) {
      uint32_t nuse[2] = {0u, 0u};
      long long a_hf = 0, a_st = 0;
      const long long a_t0 = clock64();
      const bool acct = kMlpAcct && (p.dbg & 8) && blockIdx.x == 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        for (int j = 0; j < NC; ++j) {
          const int b = j & 1;
          MLP_WAIT(smem_u32(&h_full[b]), nuse[b] & 1u, a_hf);     // all 128 rows written and fenced
          ++nuse[b];
          if (p.store_h) {
            const long long ts = kMlpAcct ? clock64() : 0;
            tma_store_2d(&tmH, hb(b), j * HC, (p.reverse ? p.tiles - 1 - tile : tile) * BM);
            tma_store_2d(&tmH, hb(b) + KB_BYTES, j * HC + 64, (p.reverse ? p.tiles - 1 - tile : tile) * BM);
            tma_store_commit();
            tma_store_wait_read<0>();
            if (kMlpAcct) a_st += clock64() - ts;
          }
          mbar_arrive(smem_u32(&h_empty[b]));
        }
      }
      tma_store_wait_all();
      if (acct) printf("mlp acct: chunk storer total %lld clk, waiting h_full %lld, store issue + read wait %lld\n", clock64() - a_t0, a_hf, a_st);
    }
  } else if (warp >= 4) {
    // ===================== epilogue warpgroups =====================
    const int wg = (warp - 4) >> 2;                          // 0: even chunks, 1: odd chunks + the output tile
    const int q = warp & 3;                                  // TMEM lane quarter
    const int wg_tid = static_cast<int>(threadIdx.x) - 128 - wg * 128;
    const int r = q * 32 + lane;                             // row within the tile
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t t_acc1 = t_lane + static_cast<uint32_t>(wg * HC);
    const uint32_t hbuf = hb(wg);
    const uint32_t hrow = hbuf + r * 128;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    const uint32_t bias1 = smem_base + OFF_BIAS1 + wg * (HC * 4);
    const uint32_t bias2 = smem_base + OFF_BIAS2;
    uint32_t nuse = 0;                                       // chunks this warpgroup has processed
    uint32_t t_par = 0;
    long long a_a1 = 0, a_hin = 0, a_a2 = 0, a_math = 0, a_out = 0;
    const long long a_t0 = clock64();
    const bool acct = kMlpAcct && (p.dbg & 8) && blockIdx.x == 0 && q == 0 && lane == 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, t_par ^= 1u) {
      const int64_t grow = static_cast<int64_t>(p.reverse ? p.tiles - 1 - tile : tile) * BM + r;
      const bool row_ok = grow < p.M;
      for (int j = wg; j < NC; j += 2, ++nuse) {
        const uint32_t par = nuse & 1u;
        float bias_v = 0.f;
        if (!BWD && p.b1) bias_v = __ldg(p.b1 + j * HC + wg_tid);
        MLP_WAIT(smem_u32(&acc1_full[wg]), par, a_a1);
        if (!BWD) MLP_WAIT(smem_u32(&h_empty[wg]), par ^ 1u, a_hin);     // operand buffer free (GEMM2 + store of chunk j-2)
        else MLP_WAIT(smem_u32(&hin_full[wg]), par, a_hin);              // hpre chunk landed in the operand buffer
        tc_fence_after();
        const long long tm0 = kMlpAcct ? clock64() : 0;
        if (!BWD) {
          named_bar_sync(1 + wg, 128);                                   // previous chunk's bias reads are done
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias1 + wg_tid * 4), "f"(bias_v) : "memory");
          named_bar_sync(1 + wg, 128);
        }
        bf16* hpre_row = (!BWD && p.hpre && row_ok) ? p.hpre + grow * p.R + j * HC : nullptr;

        uint32_t v[2][16];
        float4 bv[2][4];
        uint4 iv[2][2];
        // h4 (the 16-column group inside a 64-column half) is a compile-time value in the unrolled loop below: the
        // swizzled piece offsets are an XOR of swz (sw << 4, the low bits of the row address) with a constant
        const uint32_t swz_blk = hrow | (sw << 4);
        auto issue = [&](int s, int h4, uint32_t (&vv)[16], float4 (&bb)[4], uint4 (&ii)[2]) {
          tmem_ld_32x32b_x16(t_acc1 + s * 16, vv);
          const uint32_t blk = swz_blk + (s >> 2) * KB_BYTES;
          if (!BWD) {
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) bb[qd] = lds_f4(bias1 + s * 64 + qd * 16);
          } else {
            ii[0] = ld_shared_v4(blk ^ (32u * h4));
            ii[1] = ld_shared_v4(blk ^ (32u * h4 + 16u));
          }
        };
        auto process = [&](int s, int h4, const uint32_t (&vv)[16], const float4 (&bb)[4], const uint4 (&ii)[2]) {
          const uint32_t blk = swz_blk + (s >> 2) * KB_BYTES;
          const uint32_t adr_a = blk ^ (32u * h4), adr_b = blk ^ (32u * h4 + 16u);
          uint64_t x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = pk2u(vv[2 * i], vv[2 * i + 1]);
          uint32_t o[8];
          if (!BWD) {
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
              x[2 * qd] = add2(x[2 * qd], pk2(bb[qd].x, bb[qd].y));
              x[2 * qd + 1] = add2(x[2 * qd + 1], pk2(bb[qd].z, bb[qd].w));
            }
            if (hpre_row) {
              uint32_t a[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) a[i] = pack2_bf16(x[i]);
              st_global_v4(hpre_row + s * 16, a[0], a[1], a[2], a[3]);
              st_global_v4(hpre_row + s * 16 + 8, a[4], a[5], a[6], a[7]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = pack2_bf16(gelu_fwd2(x[i]));
          } else {
            const uint32_t iw[8] = {ii[0].x, ii[0].y, ii[0].z, ii[0].w, ii[1].x, ii[1].y, ii[1].z, ii[1].w};
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = pack2_bf16(mul2(x[i], gelu_bwd2(unpack2_bf16(iw[i]))));
          }
          st_shared_v4(adr_a, o[0], o[1], o[2], o[3]);
          st_shared_v4(adr_b, o[4], o[5], o[6], o[7]);
        };
        issue(0, 0, v[0], bv[0], iv[0]);
#pragma unroll 1
        for (int s = 0; s < HC / 16; s += 4) {               // one 64-column half per iteration, four steps unrolled
          tmem_ld_wait();
          issue(s + 1, 1, v[1], bv[1], iv[1]);
          process(s, 0, v[0], bv[0], iv[0]);
          tmem_ld_wait();
          issue(s + 2, 2, v[0], bv[0], iv[0]);
          process(s + 1, 1, v[1], bv[1], iv[1]);
          tmem_ld_wait();
          issue(s + 3, 3, v[1], bv[1], iv[1]);
          process(s + 2, 2, v[0], bv[0], iv[0]);
          tmem_ld_wait();
          if (s + 4 < HC / 16) issue(s + 4, 0, v[0], bv[0], iv[0]);
          process(s + 3, 3, v[1], bv[1], iv[1]);
        }
        tc_fence_before();
        mbar_arrive(smem_u32(&acc1_empty[wg]));              // accumulator drained
        fence_proxy_async_smem();                            // operand buffer writes -> visible to UMMA / TMA
        mbar_arrive(smem_u32(&h_full[wg]));
        if (kMlpAcct) a_math += clock64() - tm0;
      }
      {
        // ---- output tile: acc2 (+ b2 + residual) -> bf16, 32 bytes per row per step straight to global.  Each
        //      warpgroup takes 128 of the 256 columns (warpgroup 0 would otherwise idle from its last chunk until
        //      the next tile's first GEMM1) ----
        MLP_WAIT(smem_u32(&acc2_full), t_par, a_a2);
        tc_fence_after();
        const long long to0 = kMlpAcct ? clock64() : 0;
        bf16* out_row = p.out + grow * DM;
        const bf16* res_row = (!BWD && p.resid && row_ok) ? p.resid + grow * DM : nullptr;
        uint32_t v[2][16];
        float4 bv[2][4];
        uint4 iv[2][2];
        auto issue = [&](int s, uint32_t (&vv)[16], float4 (&bb)[4], uint4 (&ii)[2]) {
          tmem_ld_32x32b_x16(t_lane + ACC2_COL + s * 16, vv);
          if (!BWD) {
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) bb[qd] = lds_f4(bias2 + s * 64 + qd * 16);
            if (res_row) {
              ii[0] = ld_global_nc_v4(res_row + s * 16);
              ii[1] = ld_global_nc_v4(res_row + s * 16 + 8);
            } else {
              ii[0] = make_uint4(0u, 0u, 0u, 0u);
              ii[1] = ii[0];
            }
          }
        };
        // fused LayerNorm (forward, p.ln_out): pass 1 below also accumulates sum x / sum x^2 of the bf16-rounded values of
        // this warpgroup's 128 columns; the partial sums of a row are exchanged with the partner warpgroup through the
        // first 2 KB of the hidden-chunk operand buffers (free between the tile's last GEMM2 — acc2_full — and the next
        // tile's first chunk epilogue, which every warpgroup only starts after its own pass 2); pass 2 re-reads the
        // accumulator from TMEM (it is only released afterwards), redoes + b2 + residual and writes LN(x).
        const bool do_ln = !BWD && p.ln_out != nullptr;
        uint64_t s1 = 0ull, s2 = 0ull;                        // packed (0.f, 0.f)
        if (do_ln) {
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(hbuf + wg_tid * 4), "f"(__ldg(p.ln_gamma + wg * 128 + wg_tid)) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(hbuf + 512 + wg_tid * 4), "f"(__ldg(p.ln_beta + wg * 128 + wg_tid)) : "memory");
        }
        auto process = [&](int s, const uint32_t (&vv)[16], const float4 (&bb)[4], const uint4 (&ii)[2]) {
          uint64_t x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = pk2u(vv[2 * i], vv[2 * i + 1]);
          if (!BWD) {
            const uint32_t iw[8] = {ii[0].x, ii[0].y, ii[0].z, ii[0].w, ii[1].x, ii[1].y, ii[1].z, ii[1].w};
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
              x[2 * qd] = add2(add2(x[2 * qd], pk2(bb[qd].x, bb[qd].y)), unpack2_bf16(iw[2 * qd]));
              x[2 * qd + 1] = add2(add2(x[2 * qd + 1], pk2(bb[qd].z, bb[qd].w)), unpack2_bf16(iw[2 * qd + 1]));
            }
          }
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = pack2_bf16(x[i]);
          if (do_ln) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint64_t xr = unpack2_bf16(o[i]);
              s1 = add2(s1, xr);
              s2 = fma2(xr, xr, s2);
            }
          }
          if (row_ok) {
            st_global_v4(out_row + s * 16, o[0], o[1], o[2], o[3]);
            st_global_v4(out_row + s * 16 + 8, o[4], o[5], o[6], o[7]);
          }
        };
        const int s_lo = wg * (DM / 32), s_hi = s_lo + DM / 32;
        issue(s_lo, v[0], bv[0], iv[0]);
#pragma unroll 1
        for (int s = s_lo; s < s_hi; s += 2) {
          tmem_ld_wait();
          issue(s + 1, v[1], bv[1], iv[1]);
          process(s, v[0], bv[0], iv[0]);
          tmem_ld_wait();
          if (s + 2 < s_hi) issue(s + 2, v[0], bv[0], iv[0]);
          process(s + 1, v[1], bv[1], iv[1]);
        }
        if (do_ln) {
          float a0, a1, q0, q1;
          upk2(s1, a0, a1);
          upk2(s2, q0, q1);
          // my partial sums into the PARTNER's buffer, the partner's arrive in mine
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(hb(wg ^ 1) + 1024 + r * 8), "f"(a0 + a1), "f"(q0 + q1) : "memory");
          named_bar_sync(3, 256);                             // both warpgroups: partial sums and gamma / beta are in place
          float pa, pq;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(pa), "=f"(pq) : "r"(hbuf + 1024 + r * 8));
          const float mean = (a0 + a1 + pa) * (1.0f / DM);
          const float var = fmaxf((q0 + q1 + pq) * (1.0f / DM) - mean * mean, 0.0f);
          const float rstd = rsqrtf(var + p.ln_eps);
          if (row_ok && wg == 0) {
            if (p.ln_mean) p.ln_mean[grow] = mean;
            if (p.ln_rstd) p.ln_rstd[grow] = rstd;
          }
          const uint64_t rs2 = pk2(rstd, rstd), nm2 = pk2(-mean * rstd, -mean * rstd);
          bf16* ln_row = p.ln_out + grow * DM;
          auto pass2 = [&](int s, const uint32_t (&vv)[16], const float4 (&bb)[4], const uint4 (&ii)[2]) {
            const uint32_t iw[8] = {ii[0].x, ii[0].y, ii[0].z, ii[0].w, ii[1].x, ii[1].y, ii[1].z, ii[1].w};
            uint32_t o[8];
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
              const int cl = (s - s_lo) * 16 + qd * 4;        // column inside this warpgroup's half
              const float4 g = lds_f4(hbuf + cl * 4), bt = lds_f4(hbuf + 512 + cl * 4);
              uint64_t x0 = add2(add2(pk2u(vv[4 * qd], vv[4 * qd + 1]), pk2(bb[qd].x, bb[qd].y)), unpack2_bf16(iw[2 * qd]));
              uint64_t x1 = add2(add2(pk2u(vv[4 * qd + 2], vv[4 * qd + 3]), pk2(bb[qd].z, bb[qd].w)), unpack2_bf16(iw[2 * qd + 1]));
              x0 = unpack2_bf16(pack2_bf16(x0));              // the normalised value is LN of the bf16 x that was stored
              x1 = unpack2_bf16(pack2_bf16(x1));
              o[2 * qd] = pack2_bf16(fma2(fma2(x0, rs2, nm2), pk2(g.x, g.y), pk2(bt.x, bt.y)));
              o[2 * qd + 1] = pack2_bf16(fma2(fma2(x1, rs2, nm2), pk2(g.z, g.w), pk2(bt.z, bt.w)));
            }
            if (row_ok) {
              st_global_v4(ln_row + s * 16, o[0], o[1], o[2], o[3]);
              st_global_v4(ln_row + s * 16 + 8, o[4], o[5], o[6], o[7]);
            }
          };
          issue(s_lo, v[0], bv[0], iv[0]);
#pragma unroll 1
          for (int s = s_lo; s < s_hi; s += 2) {
            tmem_ld_wait();
            issue(s + 1, v[1], bv[1], iv[1]);
            pass2(s, v[0], bv[0], iv[0]);
            tmem_ld_wait();
            if (s + 2 < s_hi) issue(s + 2, v[0], bv[0], iv[0]);
            pass2(s + 1, v[1], bv[1], iv[1]);
          }
        }
        tc_fence_before();
        mbar_arrive(smem_u32(&acc2_empty));
        if (kMlpAcct) a_out += clock64() - to0;
      }
    }
    if (acct)
      printf("mlp acct: epilogue wg %d total %lld clk, waiting acc1_full %lld, %s %lld, acc2_full %lld; chunk math %lld, output tile %lld\n",
             wg, clock64() - a_t0, a_a1, BWD ? "hin_full (hpre landed)" : "h_empty (buffer free)", a_hin, a_a2, a_math, a_out);
  }

  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();                               // the peer may still multicast commits into this CTA
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

int g_mlp_dbg = 0;

template <int BWD, int MC>
int launch_mlp_impl(const CUtensorMap& tmX, const CUtensorMap& tmWa, const CUtensorMap& tmWb, const CUtensorMap& tmH,
                    const CUtensorMap& tmHin, const CUtensorMap& tmResid, const MlpParams& p, cudaStream_t st,
                    const char* name) {
  static bool attr_set_dev[64] = {};                         // the opt-in is a per-device attribute of the function
  const int dev = sc_device_index();
  if (dev < 0) { sc_set_error("%s: cudaGetDevice failed", name); return SC_E_LAUNCH; }
  bool& attr_set = attr_set_dev[dev];
  auto kern = mlp_fused_kernel<BWD, MC>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
    if (e != cudaSuccess) {
      sc_set_error("%s: cannot set %d B dynamic smem: %s", name, SMEM_TOTAL, cudaGetErrorString(e));
      return SC_E_LAUNCH;
    }
    attr_set = true;
  }
  int grid = sc_num_sms();
  if (p.tiles < grid) grid = p.tiles;
  if (MC) grid &= ~1;                 // whole pairs; tiles is even, so both CTAs of a pair walk the same number of tiles
  cudaError_t e = sc_launch_kernel_cluster(kern, MC ? 2 : 1, dim3(grid), dim3(THREADS), SMEM_TOTAL, st, tmX, tmWa, tmWb,
                                           tmH, tmHin, tmResid, p);
  sc_count_launch(1);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    sc_set_error("%s: launch failed: %s", name, cudaGetErrorString(e));
    return SC_E_LAUNCH;
  }
  return SC_OK;
}

template <int BWD>
int launch_mlp(const CUtensorMap& tmX, const CUtensorMap& tmWa, const CUtensorMap& tmWb, const CUtensorMap& tmH,
               const CUtensorMap& tmHin, const CUtensorMap& tmResid, const MlpParams& p, cudaStream_t st,
               const char* name) {
  // backward: weight stages shared by CTA pairs (multicast) whenever the tiles split evenly; mlp_dbg bit 4 = off
  if (BWD && p.tiles >= 2 && p.tiles % 2 == 0 && !(g_mlp_dbg & 4))
    return launch_mlp_impl<BWD, BWD ? 1 : 0>(tmX, tmWa, tmWb, tmH, tmHin, tmResid, p, st, name);
  return launch_mlp_impl<BWD, 0>(tmX, tmWa, tmWb, tmH, tmHin, tmResid, p, st, name);
}

int check_common(const char* name, int64_t M, int D, int R) {
  SC_REQUIRE(M > 0 && M < (1LL << 31) - BM, SC_E_SHAPE, "%s: bad M=%lld", name, (long long)M);
  SC_REQUIRE(D == DM, SC_E_SHAPE, "%s: model width must be %d (got %d)", name, DM, D);
  SC_REQUIRE(R >= HC && R % HC == 0, SC_E_SHAPE, "%s: hidden width must be a multiple of %d (got %d)", name, HC, R);
  return SC_OK;
}

}  // namespace

static int mlp_fwd_impl(const void* x, const void* w1, const float* b1, const void* w2, const float* b2,
                        const void* resid, void* hpre, void* h, void* out, const float* ln_gamma, const float* ln_beta,
                        float ln_eps, void* ln_out, float* ln_mean, float* ln_rstd, int64_t M, int D, int R, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(x && w1 && w2 && out, SC_E_SHAPE, "sc_mlp_fwd: null operand");
  SC_REQUIRE(!ln_out || (ln_gamma && ln_beta && !hpre && !h && sc_aligned16(ln_out)), SC_E_SHAPE,
             "sc_mlp_fwd_ln: needs gamma and beta, a 16-byte aligned ln_out, and the inference form (no hpre / h saves)");
  if ((rc = check_common("sc_mlp_fwd", M, D, R)) != SC_OK) return rc;
  SC_REQUIRE(sc_aligned16(x) && sc_aligned16(w1) && sc_aligned16(w2) && sc_aligned16(resid) && sc_aligned16(hpre) &&
                 sc_aligned16(h) && sc_aligned16(out) && sc_aligned16(b1) && sc_aligned16(b2),
             SC_E_ALIGN, "sc_mlp_fwd: all pointers must be 16-byte aligned");
  MlpParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.R = R; p.tiles = static_cast<int>((M + BM - 1) / BM);
  p.b1 = b1; p.b2 = b2;
  p.resid = static_cast<const bf16*>(resid);
  p.hpre = static_cast<bf16*>(hpre);
  p.out = static_cast<bf16*>(out);
  p.store_h = h != nullptr;
  p.ln_gamma = ln_gamma; p.ln_beta = ln_beta; p.ln_eps = ln_eps;
  p.ln_out = static_cast<bf16*>(ln_out); p.ln_mean = ln_mean; p.ln_rstd = ln_rstd;
  p.dbg = g_mlp_dbg;
  p.reverse = sc_take_direction();
  CUtensorMap tmX, tmWa, tmWb, tmH, tmHin, tmResid;
  memset(&tmH, 0, sizeof(tmH)); memset(&tmHin, 0, sizeof(tmHin)); memset(&tmResid, 0, sizeof(tmResid));
  if ((rc = sc_encode_tmap_2d_bf16(&tmX, x, DM, (uint64_t)M, DM * 2, 64, BM)) != SC_OK) return rc;
  if ((rc = sc_encode_tmap_2d_bf16(&tmWa, w1, DM, (uint64_t)R, DM * 2, 64, HC)) != SC_OK) return rc;           // [R, 256]
  if ((rc = sc_encode_tmap_2d_bf16(&tmWb, w2, (uint64_t)R, DM, (uint64_t)R * 2, 64, DM)) != SC_OK) return rc;  // [256, R]
  if (h && (rc = sc_encode_tmap_2d_bf16(&tmH, h, (uint64_t)R, (uint64_t)M, (uint64_t)R * 2, 64, BM)) != SC_OK) return rc;
  if (resid && (rc = sc_encode_tmap_2d_bf16(&tmResid, resid, DM, (uint64_t)M, DM * 2, 64, BM)) != SC_OK) return rc;
  return launch_mlp<0>(tmX, tmWa, tmWb, tmH, tmHin, tmResid, p, static_cast<cudaStream_t>(stream), "sc_mlp_fwd");
}

extern "C" int sc_mlp_fwd(const void* x, const void* w1, const float* b1, const void* w2, const float* b2,
                          const void* resid, void* hpre, void* h, void* out, int64_t M, int D, int R, void* stream) {
  return mlp_fwd_impl(x, w1, b1, w2, b2, resid, hpre, h, out, nullptr, nullptr, 0.f, nullptr, nullptr, nullptr, M, D, R,
                      stream);
}

extern "C" int sc_mlp_fwd_ln(const void* x, const void* w1, const float* b1, const void* w2, const float* b2,
                             const void* resid, void* out, const float* ln_gamma, const float* ln_beta, float ln_eps,
                             void* ln_out, float* ln_mean, float* ln_rstd, int64_t M, int D, int R, void* stream) {
  SC_REQUIRE(ln_out, SC_E_SHAPE, "sc_mlp_fwd_ln: ln_out is NULL (use sc_mlp_fwd)");
  return mlp_fwd_impl(x, w1, b1, w2, b2, resid, nullptr, nullptr, out, ln_gamma, ln_beta, ln_eps, ln_out, ln_mean, ln_rstd,
                      M, D, R, stream);
}

extern "C" int sc_mlp_bwd_dgrad(const void* dy, const void* w1, const void* w2, const void* hpre, void* dhpre, void* dx,
                                int64_t M, int D, int R, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(dy && w1 && w2 && hpre && dhpre && dx, SC_E_SHAPE, "sc_mlp_bwd_dgrad: null operand");
  if ((rc = check_common("sc_mlp_bwd_dgrad", M, D, R)) != SC_OK) return rc;
  SC_REQUIRE(sc_aligned16(dy) && sc_aligned16(w1) && sc_aligned16(w2) && sc_aligned16(hpre) && sc_aligned16(dhpre) &&
                 sc_aligned16(dx),
             SC_E_ALIGN, "sc_mlp_bwd_dgrad: all pointers must be 16-byte aligned");
  // mlp_dbg bit 32 (opt-in): the CTA-pair kernel (cta_group::2, mlp_bwd_pair.cu) — the leader issues every MMA once for both
  // 128-row tiles, each CTA stages half of every weight tile, three hidden buffers.  Bit-identical results, but 6 % SLOWER
  // (388 vs 367 us per 262144 tokens, profiles/r2_mlp_wait_accounting.txt): with the MMA warp no longer pacing the kernel the
  // epilogue warps' chunk math takes twice as long (they contend with MMAs that now run back to back).
  if ((g_mlp_dbg & 32) && M > BM)
    return sc_mlp_bwd_pair_launch(dy, w1, w2, hpre, dhpre, dx, M, R, g_mlp_dbg, static_cast<cudaStream_t>(stream));
  MlpParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.R = R; p.tiles = static_cast<int>((M + BM - 1) / BM);
  p.out = static_cast<bf16*>(dx);
  p.store_h = 1;
  p.dbg = g_mlp_dbg;
  p.reverse = sc_take_direction();
  CUtensorMap tmX, tmWa, tmWb, tmH, tmHin, tmResid;
  memset(&tmResid, 0, sizeof(tmResid));
  if ((rc = sc_encode_tmap_2d_bf16(&tmX, dy, DM, (uint64_t)M, DM * 2, 64, BM)) != SC_OK) return rc;
  // GEMM1: dh = dy W2, B(n = hidden, k = d) is W2 [256, R] read MN-major (inner = hidden, outer = d)
  if ((rc = sc_encode_tmap_2d_bf16(&tmWa, w2, (uint64_t)R, DM, (uint64_t)R * 2, 64, 64)) != SC_OK) return rc;
  // GEMM2: dx = dhpre W1, B(n = d, k = hidden) is W1 [R, 256] read MN-major (inner = d, outer = hidden)
  if ((rc = sc_encode_tmap_2d_bf16(&tmWb, w1, DM, (uint64_t)R, DM * 2, 64, 64)) != SC_OK) return rc;
  if ((rc = sc_encode_tmap_2d_bf16(&tmH, dhpre, (uint64_t)R, (uint64_t)M, (uint64_t)R * 2, 64, BM)) != SC_OK) return rc;
  if ((rc = sc_encode_tmap_2d_bf16(&tmHin, hpre, (uint64_t)R, (uint64_t)M, (uint64_t)R * 2, 64, BM)) != SC_OK) return rc;
  return launch_mlp<1>(tmX, tmWa, tmWb, tmH, tmHin, tmResid, p, static_cast<cudaStream_t>(stream), "sc_mlp_bwd_dgrad");
}

void sc_mlp_debug_set(int v) { g_mlp_dbg = v; }
