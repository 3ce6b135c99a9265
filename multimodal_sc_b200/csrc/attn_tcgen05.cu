// attn_tcgen05.cu — fused multi-head self-attention, forward and backward, for S in {64,128},
// d_head = 64 (SPEC v0 rows a1/a2/a5: the attention inside every transformer block).
//
// Work item = 128 consecutive token rows x one head.  For S = 128 that is one (sample, head);
// for S = 64 it is two samples whose score matrix is made block-diagonal by masking.  The
// whole 128x128 score tile lives in TMEM, so there is no online-softmax loop.
//
// Persistent CTAs (one per SM), 192 threads:
//   warp 0 : TMEM allocation, then lane 0 = TMA producer (Q, K, V [, dO] tiles, 2-stage ring)
//   warp 1 : lane 0 = MMA issuer (tcgen05.mma 128xNx16, accumulators in TMEM)
//   warps 2-5 : softmax / gradient math, thread = one score row (tcgen05.ld 32x32b), P / dS written
//               to shared memory in the K-major SWIZZLE_128B layout the next MMA consumes, outputs
//               staged in shared memory and written with TMA stores.
//
// forward :  S = Q K^T -> P = exp2(S c - max) (unnormalised) -> O = P V / rowsum ; lse2 = max + log2(rowsum)
// backward:  S = Q K^T, dP = dO V^T -> P = exp2(S c - lse2), delta = rowsum(P o dP),
//            dS = P o (dP - delta) / sqrt(d) -> dV = P^T dO, dK = dS^T Q, dQ = dS K
//            (delta needs no O tile: the whole row of P and dP is in one thread's registers)
#include "common.cuh"

namespace {

constexpr int kThreads = 192;
constexpr int kBwdThreads = 64 + 256;              // producer + MMA warps, two compute warpgroups
constexpr int kTile = 128 * 64 * 2;                 // one 128x64 bf16 tile = 16 KB
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void compute_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void compute_barrier256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// write 32 fp32 values (columns j0 .. j0+31 of row r) as bf16 into a 128x128 K-major SW128 operand
// made of two 64-column k-blocks of 16 KB each.
__device__ __forceinline__ void store_row_chunk_bf16(uint32_t base, int r, int j0, const float (&x)[32]) {
  const uint32_t blk = base + (j0 >> 6) * kTile;
  const int pc0 = (j0 & 63) >> 3;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    st_shared_v4(sw128(blk, r, pc0 + i), pack_bf16(x[8 * i], x[8 * i + 1]), pack_bf16(x[8 * i + 2], x[8 * i + 3]),
                 pack_bf16(x[8 * i + 4], x[8 * i + 5]), pack_bf16(x[8 * i + 6], x[8 * i + 7]));
}
__device__ __forceinline__ void store_row_chunk_zero(uint32_t base, int r, int j0) {
  const uint32_t blk = base + (j0 >> 6) * kTile;
  const int pc0 = (j0 & 63) >> 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) st_shared_v4(sw128(blk, r, pc0 + i), 0u, 0u, 0u, 0u);
}

// MMA helpers (issued by one thread).  Descriptor conventions as in gemm_tcgen05.cu.
// D[128xN] (+)= A[128 x 64, K-major tile] * B[N x 64, K-major tile]^T    (4 x UMMA_K=16)
__device__ __forceinline__ void mma_kk_64(uint32_t d_tmem, uint32_t a_tile, uint32_t b_tile, uint32_t idesc, bool acc) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    umma_bf16_ss(d_tmem, umma_smem_desc(a_tile + k * 32, 16, 1024), umma_smem_desc(b_tile + k * 32, 16, 1024), idesc,
                 (acc || k > 0) ? 1u : 0u);
}

struct FwdParams {
  float* lse;
  int64_t M;
  int S, H, num_items;
  float scale_log2;        // log2(e) / sqrt(d)
  int reverse;             // 1 = items are walked in descending order
};

// 2^x for x <= 0 (scores minus their row maximum / log-sum-exp): one MUFU.EX2.  exp2f() wraps the same instruction in a
// range check and two scalings for denormal results (FSETP + 2 FMUL per element, a third of the softmax loop); results
// below 2^-126 are flushed to zero here, far below what a bf16 probability can hold.
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
constexpr int kFwdStage = 3 * kTile;                       // Q, K, V
constexpr int fwd_smem(int stages) { return stages * kFwdStage + 2 * kTile + kTile + 1024; }   // + P (32 KB) + O staging (16 KB)

// STAGES = 1: 97 KB of shared memory and 256 TMEM columns per CTA -> two CTAs per SM; the second CTA's TMA
// loads and MMAs overlap the first one's softmax (and vice versa).  STAGES = 2: one CTA per SM with prefetch.
template <int STAGES>
__global__ void __launch_bounds__(kThreads, STAGES == 1 ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2], s_full, p_full, o_full;
  __shared__ uint32_t tmem_slot;
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t p_buf = smem + STAGES * kFwdStage;
  const uint32_t o_buf = p_buf + 2 * kTile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(&s_full), 1);
    mbar_init(smem_u32(&p_full), 128);
    mbar_init(smem_u32(&o_full), 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 0) tmem_alloc<256>(smem_u32(&tmem_slot));
  pdl_sync();                                   // no global-memory access above this line (common.cuh)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tm_s = tmem, tm_o = tmem + 128;
  const int HD = p.H * 64;

  // The two single-thread roles run their loops with the whole warp converged and let one elected lane issue
  // (elect_one_sync, common.cuh): inside `if (lane == 0)` ptxas wraps each UTCHMMA / UTMALDG in an R2UR + ELECT
  // waterfall of ~48 instructions, longer than the 66-cycle MMAs issued here.
  if (warp == 0) {
    int n = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++n) {
      const int stage = n % STAGES;
      const uint32_t ph = (n / STAGES) & 1;
      const int ri = p.reverse ? p.num_items - 1 - item : item;
      const int h = ri % p.H, row0 = (ri / p.H) * 128;
      const uint32_t fb = smem_u32(&full_bar[stage]);
      mbar_wait(smem_u32(&empty_bar[stage]), ph ^ 1u);
      const uint32_t dst = smem + stage * kFwdStage;
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(fb, kFwdStage);
        tma_load_2d(dst, &tmQKV, fb, h * 64, row0);
        tma_load_2d(dst + kTile, &tmQKV, fb, HD + h * 64, row0);
        tma_load_2d(dst + 2 * kTile, &tmQKV, fb, 2 * HD + h * 64, row0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
    int n = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++n) {
      const int stage = n % STAGES;
      const uint32_t st = smem + stage * kFwdStage;
      if (n == 0) {
        mbar_wait(smem_u32(&full_bar[0]), 0);
        tc_fence_after();
        if (elect_one_sync()) {
          mma_kk_64(tm_s, st, st + kTile, idesc_s, false);
          umma_commit(smem_u32(&s_full));
        }
        __syncwarp();
      }
      mbar_wait(smem_u32(&p_full), n & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        // O = P V : A = P (K-major, two 64-key blocks), B = V tile [keys x d] as MN-major (N = d)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16_ss(tm_o, umma_smem_desc(p_buf + (k >> 2) * kTile + (k & 3) * 32, 16, 1024),
                       umma_smem_desc(st + 2 * kTile + k * 2048, 8192, 1024), idesc_o, k > 0 ? 1u : 0u);
        umma_commit(smem_u32(&o_full));
        umma_commit(smem_u32(&empty_bar[stage]));
      }
      __syncwarp();
      if (item + static_cast<int>(gridDim.x) < p.num_items) {
        const int ns = (n + 1) % STAGES;
        const uint32_t nst = smem + ns * kFwdStage;
        mbar_wait(smem_u32(&full_bar[ns]), ((n + 1) / STAGES) & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          mma_kk_64(tm_s, nst, nst + kTile, idesc_s, false);
          umma_commit(smem_u32(&s_full));
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const bool two_seq = (p.S == 64);
    const int c_lo = two_seq ? (r >> 6) * 2 : 0;             // first 32-column chunk this row attends to
    const int c_hi = two_seq ? c_lo + 2 : 4;
    int n = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++n) {
      const int ri = p.reverse ? p.num_items - 1 - item : item;
      const int h = ri % p.H, row0 = (ri / p.H) * 128;
      mbar_wait(smem_u32(&s_full), n & 1);
      __syncwarp();
      tc_fence_after();
      float mx = -INFINITY;
      for (int c = c_lo; c < c_hi; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tm_s + lane_off + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
      const float m2 = mx * p.scale_log2;
      float sum = 0.f;
      for (int c = 0; c < 4; ++c) {
        if (c < c_lo || c >= c_hi) {
          store_row_chunk_zero(p_buf, r, c * 32);
          continue;
        }
        uint32_t v[32];
        float x[32];
        tmem_ld_32x32b_x32(tm_s + lane_off + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          x[j] = ex2_fast(fmaf(__uint_as_float(v[j]), p.scale_log2, -m2));
          sum += x[j];
        }
        store_row_chunk_bf16(p_buf, r, c * 32, x);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(smem_u32(&p_full));
      if (row0 + r < p.M) p.lse[static_cast<int64_t>(row0 + r) * p.H + h] = m2 + log2f(sum);
      const float inv = 1.0f / sum;
      mbar_wait(smem_u32(&o_full), n & 1);
      __syncwarp();
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32b_x32(tm_o + lane_off, o0);
      tmem_ld_32x32b_x32(tm_o + lane_off + 32, o1);
      tmem_ld_wait();
      // the previous item's TMA store must have read o_buf: the issuing thread waits for it HERE, a whole softmax after
      // it issued the store (waiting right after the issue put the read latency on that warp in every item)
      if (threadIdx.x == 64) tma_store_wait_read<0>();
      compute_barrier();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        st_shared_v4(sw128(o_buf, r, i),
                     pack_bf16(__uint_as_float(o0[8 * i]) * inv, __uint_as_float(o0[8 * i + 1]) * inv),
                     pack_bf16(__uint_as_float(o0[8 * i + 2]) * inv, __uint_as_float(o0[8 * i + 3]) * inv),
                     pack_bf16(__uint_as_float(o0[8 * i + 4]) * inv, __uint_as_float(o0[8 * i + 5]) * inv),
                     pack_bf16(__uint_as_float(o0[8 * i + 6]) * inv, __uint_as_float(o0[8 * i + 7]) * inv));
        st_shared_v4(sw128(o_buf, r, 4 + i),
                     pack_bf16(__uint_as_float(o1[8 * i]) * inv, __uint_as_float(o1[8 * i + 1]) * inv),
                     pack_bf16(__uint_as_float(o1[8 * i + 2]) * inv, __uint_as_float(o1[8 * i + 3]) * inv),
                     pack_bf16(__uint_as_float(o1[8 * i + 4]) * inv, __uint_as_float(o1[8 * i + 5]) * inv),
                     pack_bf16(__uint_as_float(o1[8 * i + 6]) * inv, __uint_as_float(o1[8 * i + 7]) * inv));
      }
      fence_proxy_async_smem();
      compute_barrier();
      if (threadIdx.x == 64) {
        tma_store_2d(&tmO, o_buf, h * 64, row0);
        tma_store_commit();
      }
    }
    if (threadIdx.x == 64) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<256>(tmem);
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
struct BwdParams {
  const float* lse;
  int64_t M;
  int S, H, num_items;
  float scale_log2;        // log2(e) / sqrt(d)
  float scale;             // 1 / sqrt(d)
  int reverse;             // 1 = items are walked in descending order
};

constexpr int kBwdStage = 4 * kTile;                       // Q, K, V, dO
constexpr int kBwdSmem = 2 * kBwdStage + 4 * kTile + 1024; // + P (32 KB) + dS (32 KB)

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmDQKV, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[2], drained_bar[2], sdp_full, ds_full, grads_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float delta_part[2][128];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t p_buf = smem + 2 * kBwdStage;
  const uint32_t ds_buf = p_buf + 2 * kTile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&drained_bar[s]), 256); }
    mbar_init(smem_u32(&sdp_full), 1);
    mbar_init(smem_u32(&ds_full), 256);
    mbar_init(smem_u32(&grads_full), 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQKV);
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
  pdl_sync();                                   // no global-memory access above this line (common.cuh)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tm_s = tmem, tm_dp = tmem + 128, tm_dv = tmem + 256, tm_dk = tmem + 320, tm_dq = tmem + 384;
  const int HD = p.H * 64;

  // single-thread roles: converged warp + one elected lane (see the forward kernel)
  if (warp == 0) {
    // Loads AND gradient stores.  The compute warpgroups stage dQ / dK / dV of item n in the Q / K / V slots of the
    // item's own load stage and arrive on drained_bar[stage]; this warp then stores them, waits until the stores have
    // read shared memory, and only then loads item n + 2 into the stage.  The store latency sits on this warp, one
    // item of slack ahead of the MMA warp, instead of in front of the next item's softmax.
    auto coords = [&](int it, int& h, int& row0) {
      const int ri = p.reverse ? p.num_items - 1 - it : it;
      h = ri % p.H;
      row0 = (ri / p.H) * 128;
    };
    auto store_item = [&](int k) {                            // k-th item of this CTA
      int h, row0;
      coords(static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x), h, row0);
      const uint32_t stg = smem + (k & 1) * kBwdStage;
      mbar_wait(smem_u32(&drained_bar[k & 1]), (k >> 1) & 1);
      if (lane == 0) {
        tma_store_2d(&tmDQKV, stg + 2 * kTile, 2 * HD + h * 64, row0);    // dV
        tma_store_2d(&tmDQKV, stg + kTile, HD + h * 64, row0);            // dK
        tma_store_2d(&tmDQKV, stg, h * 64, row0);                         // dQ
        tma_store_commit();
        tma_store_wait_read<0>();
      }
      __syncwarp();
    };
    int n = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++n) {
      const int stage = n & 1;
      int h, row0;
      coords(item, h, row0);
      const uint32_t fb = smem_u32(&full_bar[stage]);
      if (n >= 2) store_item(n - 2);                          // frees the stage
      const uint32_t dst = smem + stage * kBwdStage;
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(fb, kBwdStage);
        tma_load_2d(dst, &tmQKV, fb, h * 64, row0);
        tma_load_2d(dst + kTile, &tmQKV, fb, HD + h * 64, row0);
        tma_load_2d(dst + 2 * kTile, &tmQKV, fb, 2 * HD + h * 64, row0);
        tma_load_2d(dst + 3 * kTile, &tmDO, fb, h * 64, row0);
      }
      __syncwarp();
    }
    for (int k = n >= 2 ? n - 2 : 0; k < n; ++k) store_item(k);
    if (lane == 0) tma_store_wait_all();
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_tt = umma_idesc_bf16(128, 64, 1, 1);     // A^T from P/dS (MN-major), B MN-major
    constexpr uint32_t idesc_kt = umma_idesc_bf16(128, 64, 0, 1);     // A K-major, B MN-major
    int n = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++n) {
      const int stage = n & 1;
      const uint32_t st = smem + stage * kBwdStage;
      const uint32_t sQ = st, sK = st + kTile, sV = st + 2 * kTile, sDO = st + 3 * kTile;
      if (n == 0) {
        mbar_wait(smem_u32(&full_bar[0]), 0);
        tc_fence_after();
        if (elect_one_sync()) {
          mma_kk_64(tm_s, sQ, sK, idesc_s, false);            // S  = Q K^T
          mma_kk_64(tm_dp, sDO, sV, idesc_s, false);          // dP = dO V^T
          umma_commit(smem_u32(&sdp_full));
        }
        __syncwarp();
      }
      mbar_wait(smem_u32(&ds_full), n & 1);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // dV[key, d] = sum_q P[q, key] dO[q, d] : A = P^T (MN-major: M = keys, 2 chunks of 64, K = queries)
          umma_bf16_ss(tm_dv, umma_smem_desc(p_buf + k * 2048, kTile, 1024), umma_smem_desc(sDO + k * 2048, 8192, 1024),
                       idesc_tt, k > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // dK[key, d] = sum_q dS[q, key] Q[q, d]
          umma_bf16_ss(tm_dk, umma_smem_desc(ds_buf + k * 2048, kTile, 1024), umma_smem_desc(sQ + k * 2048, 8192, 1024),
                       idesc_tt, k > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // dQ[q, d] = sum_key dS[q, key] K[key, d] : A = dS K-major (2 key blocks), B = K tile MN-major
          umma_bf16_ss(tm_dq, umma_smem_desc(ds_buf + (k >> 2) * kTile + (k & 3) * 32, 16, 1024),
                       umma_smem_desc(sK + k * 2048, 8192, 1024), idesc_kt, k > 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&grads_full));
        // (the stage is released by the store thread: the gradient tiles are staged in this stage's Q / K / V slots)
      }
      __syncwarp();
      if (item + static_cast<int>(gridDim.x) < p.num_items) {
        const int ns = (n + 1) & 1;
        const uint32_t nst = smem + ns * kBwdStage;
        mbar_wait(smem_u32(&full_bar[ns]), ((n + 1) >> 1) & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          mma_kk_64(tm_s, nst, nst + kTile, idesc_s, false);
          mma_kk_64(tm_dp, nst + 3 * kTile, nst + 2 * kTile, idesc_s, false);
          umma_commit(smem_u32(&sdp_full));
        }
        __syncwarp();
      }
    }
  } else {
    // two compute warpgroups: WG g owns score columns [64 g, 64 g + 64) in both passes, partial row sums of
    // P o dP are exchanged through shared memory; the gradient tiles are drained half each.
    const int q = warp & 3;
    const int g = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const bool two_seq = (p.S == 64);
    const bool active = !two_seq || (r >> 6) == g;          // S = 64: a row only attends its own sequence's 64 keys
    int n = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++n) {
      const int ri = p.reverse ? p.num_items - 1 - item : item;
      const int h = ri % p.H, row0 = (ri / p.H) * 128;
      const float lse2 = (row0 + r < p.M) ? p.lse[static_cast<int64_t>(row0 + r) * p.H + h] : 0.f;
      mbar_wait(smem_u32(&sdp_full), n & 1);
      __syncwarp();
      tc_fence_after();
      // (no barrier here: p_buf / ds_buf were last read by the previous item's gradient MMAs, and every thread has
      //  waited on grads_full before it left that item; the gradient staging no longer aliases these buffers)
      // pass 1: P (bf16 -> smem, kept packed in registers) and this warpgroup's part of delta = rowsum(P o dP)
      uint32_t pk[32];
      float delta = 0.f;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * g + cc;
        if (!active) {
          store_row_chunk_zero(p_buf, r, c * 32);
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[cc * 16 + j] = 0u;
          continue;
        }
        uint32_t sv[32], dv[32];
        tmem_ld_32x32b_x32(tm_s + lane_off + c * 32, sv);
        tmem_ld_32x32b_x32(tm_dp + lane_off + c * 32, dv);
        tmem_ld_wait();
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          x[j] = ex2_fast(fmaf(__uint_as_float(sv[j]), p.scale_log2, -lse2));
          delta = fmaf(x[j], __uint_as_float(dv[j]), delta);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[cc * 16 + j] = pack_bf16(x[2 * j], x[2 * j + 1]);
        store_row_chunk_bf16(p_buf, r, c * 32, x);
      }
      delta_part[g][r] = delta;
      compute_barrier256();
      delta = delta_part[0][r] + delta_part[1][r];
      // pass 2: dS = P o (dP - delta) / sqrt(d)
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * g + cc;
        if (!active) {
          store_row_chunk_zero(ds_buf, r, c * 32);
          continue;
        }
        uint32_t dv[32];
        tmem_ld_32x32b_x32(tm_dp + lane_off + c * 32, dv);
        tmem_ld_wait();
        float x[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 pp = unpack_bf16(pk[cc * 16 + j]);
          x[2 * j] = pp.x * (__uint_as_float(dv[2 * j]) - delta) * p.scale;
          x[2 * j + 1] = pp.y * (__uint_as_float(dv[2 * j + 1]) - delta) * p.scale;
        }
        store_row_chunk_bf16(ds_buf, r, c * 32, x);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(smem_u32(&ds_full));
      // drain: WG0 takes dV and the first 32 columns of dQ, WG1 takes dK and the last 32 columns of dQ.  Staging = the
      // V / K / Q slots of THIS item's load stage: once the gradient MMAs are complete nothing reads them any more, and
      // the stage is handed back to the producer only after the TMA stores have read it.  (Staging in p_buf / ds_buf
      // put the store's read latency and a 256-thread barrier in front of the next item's softmax.)
      mbar_wait(smem_u32(&grads_full), n & 1);
      __syncwarp();
      tc_fence_after();
      const uint32_t stg = smem + (n & 1) * kBwdStage;
      {
        const uint32_t src = (g == 0 ? tm_dv : tm_dk) + lane_off;
        const uint32_t dst = stg + (g == 0 ? 2 * kTile : kTile);
        uint32_t o0[32], o1[32];
        tmem_ld_32x32b_x32(src, o0);
        tmem_ld_32x32b_x32(src + 32, o1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          st_shared_v4(sw128(dst, r, i),
                       pack_bf16(__uint_as_float(o0[8 * i]), __uint_as_float(o0[8 * i + 1])),
                       pack_bf16(__uint_as_float(o0[8 * i + 2]), __uint_as_float(o0[8 * i + 3])),
                       pack_bf16(__uint_as_float(o0[8 * i + 4]), __uint_as_float(o0[8 * i + 5])),
                       pack_bf16(__uint_as_float(o0[8 * i + 6]), __uint_as_float(o0[8 * i + 7])));
          st_shared_v4(sw128(dst, r, 4 + i),
                       pack_bf16(__uint_as_float(o1[8 * i]), __uint_as_float(o1[8 * i + 1])),
                       pack_bf16(__uint_as_float(o1[8 * i + 2]), __uint_as_float(o1[8 * i + 3])),
                       pack_bf16(__uint_as_float(o1[8 * i + 4]), __uint_as_float(o1[8 * i + 5])),
                       pack_bf16(__uint_as_float(o1[8 * i + 6]), __uint_as_float(o1[8 * i + 7])));
        }
        uint32_t oq[32];
        tmem_ld_32x32b_x32(tm_dq + lane_off + g * 32, oq);
        tmem_ld_wait();
        const uint32_t dq_dst = stg;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          st_shared_v4(sw128(dq_dst, r, 4 * g + i),
                       pack_bf16(__uint_as_float(oq[8 * i]), __uint_as_float(oq[8 * i + 1])),
                       pack_bf16(__uint_as_float(oq[8 * i + 2]), __uint_as_float(oq[8 * i + 3])),
                       pack_bf16(__uint_as_float(oq[8 * i + 4]), __uint_as_float(oq[8 * i + 5])),
                       pack_bf16(__uint_as_float(oq[8 * i + 6]), __uint_as_float(oq[8 * i + 7])));
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(smem_u32(&drained_bar[n & 1]));             // warp 0 stores the tiles and recycles the stage
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

int g_attn_fwd_stages = 1;

int check_shape(const char* nm, int64_t B, int S, int H) {
  SC_REQUIRE(B > 0 && (S == 64 || S == 128) && H >= 1 && H <= 16, SC_E_SHAPE,
             "%s: B=%lld S=%d H=%d (S must be 64 or 128, d_head is fixed at 64)", nm, (long long)B, S, H);
  SC_REQUIRE(B * S < (1LL << 31), SC_E_SHAPE, "%s: too many tokens", nm);
  return SC_OK;
}

}  // namespace

void sc_attn_debug_set_fwd_stages(int v) { g_attn_fwd_stages = v; }

extern "C" {

int sc_attn_fwd(const void* qkv, void* out, float* lse, int64_t B, int S, int H, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(qkv && out && lse, SC_E_SHAPE, "sc_attn_fwd: null pointer");
  if ((rc = check_shape("sc_attn_fwd", B, S, H)) != SC_OK) return rc;
  SC_REQUIRE(sc_aligned16(qkv) && sc_aligned16(out), SC_E_ALIGN, "sc_attn_fwd: unaligned pointer");
  const int64_t M = B * S;
  const int HD = H * 64;
  CUtensorMap tmQKV, tmO;
  if ((rc = sc_encode_tmap_2d_bf16(&tmQKV, qkv, 3 * HD, M, (uint64_t)3 * HD * 2, 64, 128)) != SC_OK) return rc;
  if ((rc = sc_encode_tmap_2d_bf16(&tmO, out, HD, M, (uint64_t)HD * 2, 64, 128)) != SC_OK) return rc;
  FwdParams p;
  p.lse = lse; p.M = M; p.S = S; p.H = H;
  p.num_items = (int)((M + 127) / 128) * H;
  p.scale_log2 = kLog2e * 0.125f;
  p.reverse = sc_take_direction();
  static bool attr_dev[64] = {};                             // the opt-in is a per-device attribute of the function
  const int dev = sc_device_index();
  SC_REQUIRE(dev >= 0, SC_E_LAUNCH, "sc_attn_fwd: cudaGetDevice failed");
  bool& attr = attr_dev[dev];
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem(1));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem(2));
    SC_REQUIRE(e == cudaSuccess, SC_E_LAUNCH, "sc_attn_fwd: smem attribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  if (g_attn_fwd_stages == 2) {
    const int grid = p.num_items < sc_num_sms() ? p.num_items : sc_num_sms();
    sc_launch_kernel(attn_fwd_kernel<2>, dim3(grid), dim3(kThreads), fwd_smem(2), (cudaStream_t)stream, tmQKV, tmO, p);
  } else {
    const int cap = 2 * sc_num_sms();
    const int grid = p.num_items < cap ? p.num_items : cap;
    sc_launch_kernel(attn_fwd_kernel<1>, dim3(grid), dim3(kThreads), fwd_smem(1), (cudaStream_t)stream, tmQKV, tmO, p);
  }
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_attn_fwd");
  return SC_OK;
}

int sc_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int64_t B, int S,
                int H, void* stream) {
  (void)out;   // not needed: delta = rowsum(P o dP) is computed from the in-register rows
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(qkv && dout && lse && dqkv, SC_E_SHAPE, "sc_attn_bwd: null pointer");
  if ((rc = check_shape("sc_attn_bwd", B, S, H)) != SC_OK) return rc;
  SC_REQUIRE(sc_aligned16(qkv) && sc_aligned16(dout) && sc_aligned16(dqkv), SC_E_ALIGN, "sc_attn_bwd: unaligned pointer");
  const int64_t M = B * S;
  const int HD = H * 64;
  CUtensorMap tmQKV, tmDO, tmDQKV;
  if ((rc = sc_encode_tmap_2d_bf16(&tmQKV, qkv, 3 * HD, M, (uint64_t)3 * HD * 2, 64, 128)) != SC_OK) return rc;
  if ((rc = sc_encode_tmap_2d_bf16(&tmDO, dout, HD, M, (uint64_t)HD * 2, 64, 128)) != SC_OK) return rc;
  if ((rc = sc_encode_tmap_2d_bf16(&tmDQKV, dqkv, 3 * HD, M, (uint64_t)3 * HD * 2, 64, 128)) != SC_OK) return rc;
  BwdParams p;
  p.lse = lse; p.M = M; p.S = S; p.H = H;
  p.num_items = (int)((M + 127) / 128) * H;
  p.scale_log2 = kLog2e * 0.125f;
  p.scale = 0.125f;
  p.reverse = sc_take_direction();
  static bool attr_dev[64] = {};
  const int dev = sc_device_index();
  SC_REQUIRE(dev >= 0, SC_E_LAUNCH, "sc_attn_bwd: cudaGetDevice failed");
  bool& attr = attr_dev[dev];
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
    SC_REQUIRE(e == cudaSuccess, SC_E_LAUNCH, "sc_attn_bwd: smem attribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  const int grid = p.num_items < sc_num_sms() ? p.num_items : sc_num_sms();
  sc_launch_kernel(attn_bwd_kernel, dim3(grid), dim3(kBwdThreads), kBwdSmem, (cudaStream_t)stream, tmQKV, tmDO, tmDQKV, p);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_attn_bwd");
  return SC_OK;
}

}  // extern "C"
