// mlp_bwd_pair.cu — the fused MLP backward (sc_mlp_bwd_dgrad) as CTA PAIRS (cta_group::2), sm_100a.
//
//   dhpre = (dy . W2) o gelu'(hpre),   dx = dhpre . W1          (SPEC v0 rows a1/a2/a5: the two dgrad GEMMs of the block MLP)
//
// Same pipeline as mlp_fused.cu (GEMM1 chunk in TMEM -> epilogue warps transform it in registers -> bf16 operand buffer in
// shared memory -> GEMM2 + TMA store), but a cluster of two CTAs works on a 256-row super-tile: CTA r owns rows r*128..+127
// (its own dy tile, hidden buffers, hpre / dhpre traffic and epilogues) while the MMAs are issued ONCE for both by the
// leader (rank 0) as 256 x N x 16 tcgen05.mma.cta_group::2, and every CTA stages only HALF of each weight tile.
// Why (profiles/r2_mlp_wait_accounting.txt, wait accounting of the single-CTA kernel): the MMA warp paced that kernel —
// 52 % of its time went into ISSUING the 192 MMAs of a tile (the N = 128 MMAs of GEMM1 are issue-bound) and 30 % into
// waiting for weight stages (a chunk consumes four 32 KB stages, the ring holds three).  As a pair the issue cost per
// 128-row tile halves, and the same 96 KB ring holds 1.5 chunks of weights instead of 0.75.
//
// Per CTA, 384 threads (roles as in mlp_fused.cu):
//   warp 0 : weight producer  — this CTA's half of every weight stage, completion counted on the LEADER's full barrier
//   warp 1 : MMA issuer       — leader only
//   warp 2 : TMEM allocator (cta_group::2), then chunk storer (dhpre by TMA from the operand buffer)
//   warp 3 : activation producer — own dy tile (counted on the leader's barrier), own hpre chunks (local barrier)
//   warps 4-7 / 8-11 : epilogue warpgroups (even / odd hidden chunks; both drain half of the output tile)
// Barriers that the leader's MMA warp waits on collect arrivals from BOTH CTAs (remote mbarrier.arrive.release.cluster);
// everything the MMA warp signals is a multicast tcgen05.commit that arrives in both CTAs.
#include "common.cuh"

namespace {

constexpr int BM = 128;                 // rows per CTA tile (256 per pair)
constexpr int DM = 256;                 // model width (GEMM1 K, GEMM2 N)
constexpr int HC = 128;                 // hidden columns per chunk (GEMM1 N, GEMM2 K)
constexpr int KB_BYTES = BM * 64 * 2;   // one 128-row x 64-k K-major SWIZZLE_128B block: 16 KB
constexpr int XA_BYTES = 4 * KB_BYTES;  // dy tile: 4 k-blocks
constexpr int HB_BYTES = 2 * KB_BYTES;  // one hidden chunk: 2 k-blocks
constexpr int CHUNK8K = 8192;           // one 64-wide MN-major chunk (64 k-rows x 128 B)
// weight ring stage, PER CTA (half of the operand): GEMM1: 4 k-blocks x (64 k-rows x 64 hidden columns) = 4 x 8 KB;
// GEMM2: 2 k-blocks x (64 hidden rows x 128 output columns) = 2 x 16 KB
constexpr int STAGE_BYTES = 32768;
// Two weight stages and THREE hidden buffers.  As a pair the MMA warp no longer paces the kernel; what does is the cycle
// of a hidden buffer: epilogue(j) -> GEMM2(j) + store -> hpre(j') load -> epilogue(j').  With two buffers j' = j + 2 is the
// SAME warpgroup's next chunk, so the HBM latency of the hpre load (~2.5k clk) sits on its critical path; with three
// buffers the load for chunk j + 3 is issued a chunk earlier and lands while the warpgroup still works on chunk j + 1's
// neighbour.  The weights come from L2 and are consumed at one stage per ~2.5k clk: two stages (one chunk) cover that.
constexpr int STAGES = 2;
constexpr int NHB = 3;
constexpr int OFF_XA = 0;
constexpr int OFF_HB = OFF_XA + XA_BYTES;
constexpr int OFF_RING = OFF_HB + NHB * HB_BYTES;
constexpr int SMEM_TOTAL = OFF_RING + STAGES * STAGE_BYTES;     // 229376
static_assert(SMEM_TOTAL <= 232448 - 1024, "shared memory budget");
constexpr int THREADS = 384;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t ACC2_COL = 256;

struct PairParams {
  int64_t M;
  int R;
  int pairs;                 // 256-row super-tiles
  bf16* out;                 // dx [M, 256]
  int dbg;
};

#ifdef SC_MLP_ACCT
#define PW(bar, par, ctr) do { const long long _t = clock64(); mbar_wait(bar, par); (ctr) += clock64() - _t; } while (0)
#define PWC(bar, par, ctr) do { const long long _t = clock64(); mbar_wait_cluster(bar, par); (ctr) += clock64() - _t; } while (0)
constexpr bool kAcct = true;
#else
#define PW(bar, par, ctr) mbar_wait(bar, par)
#define PWC(bar, par, ctr) mbar_wait_cluster(bar, par)
constexpr bool kAcct = false;
#endif

// wait with cluster-scope acquire: the arrivals come from the peer CTA (release.cluster) and order its shared-memory /
// tensor-memory accesses before what follows here
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t it = 0;; ++it) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
    if (done) break;
    if (it > SC_MBAR_SPIN_LIMIT) {
      printf("semcom: mbarrier timeout (mlp pair, block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- packed fp32x2 math (same forms as mlp_fused.cu) ----
__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ uint64_t pk2u(uint32_t lo, uint32_t hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t tanh2(uint64_t u) { float a, b; upk2(u, a, b); return pk2(tanh_fast(a), tanh_fast(b)); }
__device__ __forceinline__ uint32_t pack2_bf16(uint64_t v) { float a, b; upk2(v, a, b); return pack_bf16(a, b); }
__device__ __forceinline__ uint64_t unpack2_bf16(uint32_t u) { return pk2u(u << 16, u & 0xFFFF0000u); }
__device__ __forceinline__ uint64_t neg2(uint64_t a) { return a ^ 0x8000000080000000ull; }
// d/dx gelu_tanh(x), nine packed operations + tanh per pair (see mlp_fused.cu)
__device__ __forceinline__ uint64_t gelu_bwd2(uint64_t x) {
  const uint64_t k0 = pk2(SC_GELU_C0, SC_GELU_C0), k01 = pk2(SC_GELU_C0 * SC_GELU_C1, SC_GELU_C0 * SC_GELU_C1);
  const uint64_t c0 = pk2(0.5f * SC_GELU_C0, 0.5f * SC_GELU_C0);
  const uint64_t c1 = pk2(1.5f * SC_GELU_C0 * SC_GELU_C1, 1.5f * SC_GELU_C0 * SC_GELU_C1);
  const uint64_t half = pk2(0.5f, 0.5f), one = pk2(1.0f, 1.0f);
  const uint64_t x2 = mul2(x, x);
  const uint64_t t = tanh2(mul2(x, fma2(x2, k01, k0)));
  const uint64_t hxdu = mul2(x, fma2(x2, c1, c0));
  const uint64_t omt2 = fma2(neg2(t), t, one);
  return fma2(hxdu, omt2, fma2(t, half, half));
}

// four 256 x N x 16 MMAs of one 64-wide k-block (cta_group::2) in ONE asm block, optionally preceded by a non-blocking test of
// `next_bar` (result returned) and followed by a multicast commit on `commit_bar` (arrives in both CTAs)
template <bool kTest, bool kCommit>
__device__ __forceinline__ uint32_t umma_kblock4_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t a_step,
                                                     uint32_t b_step, uint32_t idesc, uint32_t accumulate_first,
                                                     uint32_t commit_bar, uint32_t next_bar, uint32_t next_parity) {
  uint32_t ready = 0;
#define SC_BODY                                                                \
  "setp.ne.b32 p, %7, 0;\n\t"                                                  \
  "setp.eq.b32 t, 0, 0;\n\t"                                                   \
  "cvt.u64.u32 sa, %4;\n\t"                                                    \
  "cvt.u64.u32 sb, %5;\n\t"                                                    \
  "tcgen05.mma.cta_group::2.kind::f16 [%1], %2, %3, %6, p;\n\t"                \
  "add.u64 a, %2, sa;\n\t"                                                     \
  "add.u64 b, %3, sb;\n\t"                                                     \
  "tcgen05.mma.cta_group::2.kind::f16 [%1], a, b, %6, t;\n\t"                  \
  "add.u64 a, a, sa;\n\t"                                                      \
  "add.u64 b, b, sb;\n\t"                                                      \
  "tcgen05.mma.cta_group::2.kind::f16 [%1], a, b, %6, t;\n\t"                  \
  "add.u64 a, a, sa;\n\t"                                                      \
  "add.u64 b, b, sb;\n\t"                                                      \
  "tcgen05.mma.cta_group::2.kind::f16 [%1], a, b, %6, t;\n\t"
#define SC_OPS                                                                                                  \
  "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(a_step), "r"(b_step), "r"(idesc), "r"(accumulate_first),          \
      "r"(commit_bar), "r"(next_bar), "r"(next_parity)
#define SC_COMMIT                                                                                                \
  "{\n\t.reg .b16 cm;\n\tmov.b16 cm, 3;\n\t"                                                                    \
  "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%8], cm;\n\t}\n\t"
  if (kTest && kCommit) {
    asm volatile("{\n\t.reg .pred p, q, t;\n\t.reg .b64 a, b, sa, sb;\n\t"
                 "mbarrier.test_wait.parity.shared::cta.b64 q, [%9], %10;\n\t" SC_BODY SC_COMMIT "selp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(ready) : SC_OPS : "memory");
  } else if (kTest) {
    asm volatile("{\n\t.reg .pred p, q, t;\n\t.reg .b64 a, b, sa, sb;\n\t"
                 "mbarrier.test_wait.parity.shared::cta.b64 q, [%9], %10;\n\t" SC_BODY "selp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(ready) : SC_OPS : "memory");
  } else if (kCommit) {
    asm volatile("{\n\t.reg .pred p, t;\n\t.reg .b64 a, b, sa, sb;\n\t" SC_BODY SC_COMMIT "mov.u32 %0, 0;\n\t}"
                 : "=r"(ready) : SC_OPS : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p, t;\n\t.reg .b64 a, b, sa, sb;\n\t" SC_BODY "mov.u32 %0, 0;\n\t}"
                 : "=r"(ready) : SC_OPS : "memory");
  }
#undef SC_BODY
#undef SC_OPS
#undef SC_COMMIT
  return ready;
}

__global__ void __launch_bounds__(THREADS, 1)
mlp_bwd_pair_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWa,
                    const __grid_constant__ CUtensorMap tmWb, const __grid_constant__ CUtensorMap tmH,
                    const __grid_constant__ CUtensorMap tmHin, const PairParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // identical layout in both CTAs: multicast commits and mapa() address the peer's barrier at the same offset
  __shared__ __align__(8) uint64_t x_full, x_empty;
  __shared__ __align__(8) uint64_t w_full[STAGES], w_empty[STAGES];
  __shared__ __align__(8) uint64_t acc1_full[2], acc1_empty[2];
  __shared__ __align__(8) uint64_t h_full[NHB], h_ready[NHB], h_empty[NHB], hin_full[NHB];
  __shared__ __align__(8) uint64_t acc2_full, acc2_empty;
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("semcom: mlp_bwd_pair dynamic shared memory base 0x%x is not 1 KB aligned\n", smem_base);
    __trap();
  }
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NC = p.R / HC;
  const uint32_t rank = cluster_ctarank();                  // 0 = MMA leader
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&x_full), 1);                         // leader: its producer's arrive.expect_tx (bytes of BOTH dy tiles)
    mbar_init(smem_u32(&x_empty), 1);                        // multicast commit of the tile's last GEMM1
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&w_full[s]), 1);                    // leader: arrive.expect_tx for both halves
      mbar_init(smem_u32(&w_empty[s]), 1);                   // multicast commit
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&acc1_full[b]), 1);                 // multicast commit
      mbar_init(smem_u32(&acc1_empty[b]), 8);                // leader: the four warps of warpgroup b in both CTAs (one arrival per warp)
    }
    for (int b = 0; b < NHB; ++b) {
      mbar_init(smem_u32(&h_full[b]), 4);                    // local: this CTA's chunk is written (storer)
      mbar_init(smem_u32(&h_ready[b]), 8);                   // leader: both CTAs' chunks are written (GEMM2)
      mbar_init(smem_u32(&h_empty[b]), 2);                   // multicast commit of GEMM2 + local store has read the buffer
      mbar_init(smem_u32(&hin_full[b]), 1);                  // local hpre load
    }
    mbar_init(smem_u32(&acc2_full), 1);                      // multicast commit
    mbar_init(smem_u32(&acc2_empty), 16);                    // leader: all eight epilogue warps of both CTAs
    fence_mbar_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWa);
    tma_prefetch_desc(&tmWb);
    tma_prefetch_desc(&tmH);
    tma_prefetch_desc(&tmHin);
  }
  if (warp == 2) tmem_alloc_2sm<TMEM_COLS>(smem_u32(&tmem_base_slot));
  pdl_sync();                                               // no global-memory access above this line (common.cuh)
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                       // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const uint32_t xa = smem_base + OFF_XA;
  const uint32_t ring = smem_base + OFF_RING;
  auto hb = [&](int b) { return smem_base + OFF_HB + b * HB_BYTES; };
  const int pair_first = static_cast<int>(blockIdx.x) >> 1;
  const int pair_stride = static_cast<int>(gridDim.x) >> 1;
  const bool acct_cta = kAcct && (p.dbg & 8) && blockIdx.x == 0;

  if (warp == 0) {
    // ===================== weight producer (both CTAs: own half, leader's barrier) =====================
    int stage = 0;
    uint32_t phase = 0;
    uint32_t ready = mbar_test(smem_u32(&w_empty[0]), 1u);
    long long a_w = 0;
    const long long a_t0 = clock64();
    auto next_stage = [&]() -> uint32_t {
      if (!ready) PW(smem_u32(&w_empty[stage]), phase ^ 1u, a_w);
      const int nstage = (stage + 1 == STAGES) ? 0 : stage + 1;
      const uint32_t nphase = (stage + 1 == STAGES) ? (phase ^ 1u) : phase;
      ready = mbar_test(smem_u32(&w_empty[nstage]), nphase ^ 1u);
      return ring + stage * STAGE_BYTES;
    };
    auto advance = [&]() { if (++stage == STAGES) { stage = 0; phase ^= 1u; } };
    auto load_g1 = [&](int j) {            // one stage: this CTA's 64 hidden columns of chunk j, all four k-blocks of W2
      const uint32_t dst = next_stage();
      const uint32_t fb_local = smem_u32(&w_full[stage]);
      const uint32_t fb = mapa_shared(fb_local, 0);
      if (elect_one_sync()) {
        if (leader) mbar_arrive_expect_tx(fb_local, 2 * STAGE_BYTES);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)     // W2 [256, R] read MN-major: 64 k-rows x 64 hidden columns
          tma_load_2d_2sm(dst + kb * CHUNK8K, &tmWa, fb, j * HC + static_cast<int>(rank) * 64, kb * 64);
      }
      __syncwarp();
      advance();
    };
    auto load_g2 = [&](int j) {            // one stage: this CTA's 128 output columns, both hidden k-blocks of chunk j of W1
      const uint32_t dst = next_stage();
      const uint32_t fb_local = smem_u32(&w_full[stage]);
      const uint32_t fb = mapa_shared(fb_local, 0);
      if (elect_one_sync()) {
        if (leader) mbar_arrive_expect_tx(fb_local, 2 * STAGE_BYTES);
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int c = 0; c < 2; ++c)      // W1 [R, 256] read MN-major: 64 k-rows (hidden) x 64 output columns
            tma_load_2d_2sm(dst + t * 2 * CHUNK8K + c * CHUNK8K, &tmWb, fb, static_cast<int>(rank) * 128 + c * 64, j * HC + t * 64);
      }
      __syncwarp();
      advance();
    };
    for (int pr = pair_first; pr < p.pairs; pr += pair_stride) {
      for (int j = 0; j < NC; ++j) {
        load_g1(j);
        if (j >= 1) load_g2(j - 1);
      }
      load_g2(NC - 1);
    }
    if (acct_cta && lane == 0) printf("mlp pair acct: weight producer total %lld clk, waiting for a free stage %lld\n", clock64() - a_t0, a_w);
  } else if (warp == 3) {
    // ===================== activation producer (own rows) =====================
    if (lane == 0) {
      uint32_t t_par = 0;
      int g = 0;                                             // running chunk index of this CTA: buffer g % 3, use g / 3
      long long a_x = 0, a_h = 0;
      const long long a_t0 = clock64();
      const uint32_t xf_local = smem_u32(&x_full);
      const uint32_t xf = mapa_shared(xf_local, 0);
      for (int pr = pair_first; pr < p.pairs; pr += pair_stride, t_par ^= 1u) {
        const int row0 = (pr * 2 + static_cast<int>(rank)) * BM;
        PW(smem_u32(&x_empty), t_par ^ 1u, a_x);
        if (leader) mbar_arrive_expect_tx(xf_local, 2 * XA_BYTES);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) tma_load_2d_2sm(xa + kb * KB_BYTES, &tmX, xf, kb * 64, row0);
        for (int j = 0; j < NC; ++j, ++g) {                  // hpre chunk j -> hidden buffer g % 3 (transformed in place)
          const int b = g % NHB;
          PW(smem_u32(&h_empty[b]), (static_cast<uint32_t>(g / NHB) & 1u) ^ 1u, a_h);
          mbar_arrive_expect_tx(smem_u32(&hin_full[b]), HB_BYTES);
          tma_load_2d(hb(b), &tmHin, smem_u32(&hin_full[b]), j * HC, row0);
          tma_load_2d(hb(b) + KB_BYTES, &tmHin, smem_u32(&hin_full[b]), j * HC + 64, row0);
        }
      }
      if (acct_cta) printf("mlp pair acct: activation producer total %lld clk, waiting x_empty %lld, h_empty %lld\n", clock64() - a_t0, a_x, a_h);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(2 * BM, HC, 0, 1);
      constexpr uint32_t idesc2 = umma_idesc_bf16(2 * BM, DM, 0, 1);
      constexpr uint32_t b_kstep16 = 2048 >> 4;              // MN-major B: 16 k-rows = 2048 bytes
      const uint64_t a_desc0 = umma_smem_desc(0, 16, 1024);
      const uint64_t b_desc0 = umma_smem_desc(0, CHUNK8K, 1024);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t n1[2] = {0u, 0u};
      int g2 = 0;                                            // running chunk index of GEMM2 (hidden buffer g2 % 3)
      uint32_t t_par = 0;
      uint32_t ready = mbar_test(smem_u32(&w_full[0]), 0u);
      int nstage = 0;
      uint32_t nphase = 0;
      long long a_w = 0, a_a1 = 0, a_hf = 0, a_a2 = 0, a_x = 0;
      const long long a_t0 = clock64();
      auto acquire_stage = [&]() -> uint32_t {
        if (!ready) PW(smem_u32(&w_full[stage]), phase, a_w);
        nstage = (stage + 1 == STAGES) ? 0 : stage + 1;
        nphase = (stage + 1 == STAGES) ? (phase ^ 1u) : phase;
        return ring + stage * STAGE_BYTES;
      };
      auto advance = [&]() { stage = nstage; phase = nphase; };
      auto gemm1 = [&](int j) {
        const int b = j & 1;
        PWC(smem_u32(&acc1_empty[b]), (n1[b] & 1u) ^ 1u, a_a1);
        ++n1[b];
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(b * HC);
        const uint32_t wb = acquire_stage();
        tc_fence_after();
        uint32_t r = 0;
        if (elect_one_sync()) {
          // sixteen 256 x 128 x 16 MMAs: four k-blocks of the dy tile against this stage's four W2 chunks
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            const uint64_t ad = a_desc0 + (((xa + kb * KB_BYTES) & 0x3FFFFu) >> 4);
            const uint64_t bd = b_desc0 + (((wb + kb * CHUNK8K) & 0x3FFFFu) >> 4);
            if (kb == 0)
              r = umma_kblock4_2sm<true, false>(d_tmem, ad, bd, 2u, b_kstep16, idesc1, 0u, 0u, smem_u32(&w_full[nstage]), nphase);
            else if (kb < 3)
              umma_kblock4_2sm<false, false>(d_tmem, ad, bd, 2u, b_kstep16, idesc1, 1u, 0u, 0u, 0u);
            else
              umma_kblock4_2sm<false, true>(d_tmem, ad, bd, 2u, b_kstep16, idesc1, 1u, smem_u32(&w_empty[stage]), 0u, 0u);
          }
        }
        ready = __shfl_sync(0xffffffffu, r, 0);
        advance();
        if (elect_one_sync()) {
          umma_commit_2sm(smem_u32(&acc1_full[b]));
          if (j == NC - 1) umma_commit_2sm(smem_u32(&x_empty));           // dy tiles free once the last GEMM1 retires
        }
        __syncwarp();
      };
      auto gemm2 = [&](int j) {
        const int b = g2 % NHB;
        PWC(smem_u32(&h_ready[b]), static_cast<uint32_t>(g2 / NHB) & 1u, a_hf);
        ++g2;
        if (j == 0) PWC(smem_u32(&acc2_empty), t_par ^ 1u, a_a2);          // previous super-tile's output drained
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ACC2_COL;
        const uint32_t wb = acquire_stage();
        tc_fence_after();
        uint32_t r = 0;
        if (elect_one_sync()) {
          // eight 256 x 256 x 16 MMAs: two hidden k-blocks of the operand buffer against this stage's W1 halves
          const uint64_t ad0 = a_desc0 + ((hb(b) & 0x3FFFFu) >> 4);
          const uint64_t bd0 = b_desc0 + ((wb & 0x3FFFFu) >> 4);
          r = umma_kblock4_2sm<true, false>(d_tmem, ad0, bd0, 2u, b_kstep16, idesc2, j > 0 ? 1u : 0u, 0u,
                                            smem_u32(&w_full[nstage]), nphase);
          umma_kblock4_2sm<false, true>(d_tmem, ad0 + (KB_BYTES >> 4), bd0 + ((2 * CHUNK8K) >> 4), 2u, b_kstep16, idesc2, 1u,
                                        smem_u32(&w_empty[stage]), 0u, 0u);
        }
        ready = __shfl_sync(0xffffffffu, r, 0);
        advance();
        if (elect_one_sync()) umma_commit_2sm(smem_u32(&h_empty[b]));
        __syncwarp();
      };
      for (int pr = pair_first; pr < p.pairs; pr += pair_stride, t_par ^= 1u) {
        PW(smem_u32(&x_full), t_par, a_x);
        tc_fence_after();
        for (int j = 0; j < NC; ++j) {
          gemm1(j);
          if (j >= 1) gemm2(j - 1);
        }
        gemm2(NC - 1);
        if (elect_one_sync()) umma_commit_2sm(smem_u32(&acc2_full));
        __syncwarp();
      }
      if (acct_cta && lane == 0)
        printf("mlp pair acct: mma total %lld clk, waiting weights %lld, acc1_empty %lld, h_ready %lld, acc2_empty %lld, x_full %lld\n",
               clock64() - a_t0, a_w, a_a1, a_hf, a_a2, a_x);
    }
  } else if (warp == 2) {
    // ===================== chunk storer (own rows) =====================
    if (lane == 0) {
      int g = 0;
      for (int pr = pair_first; pr < p.pairs; pr += pair_stride) {
        const int row0 = (pr * 2 + static_cast<int>(rank)) * BM;
        for (int j = 0; j < NC; ++j, ++g) {
          const int b = g % NHB;
          mbar_wait(smem_u32(&h_full[b]), static_cast<uint32_t>(g / NHB) & 1u);     // this CTA's 128 rows written and fenced
          tma_store_2d(&tmH, hb(b), j * HC, row0);
          tma_store_2d(&tmH, hb(b) + KB_BYTES, j * HC + 64, row0);
          tma_store_commit();
          tma_store_wait_read<0>();
          mbar_arrive(smem_u32(&h_empty[b]));
        }
      }
      tma_store_wait_all();
    }
  } else if (warp >= 4) {
    // ===================== epilogue warpgroups (own rows; arrivals for the MMA warp go to the leader) =====================
    const int wg = (warp - 4) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t t_acc1 = t_lane + static_cast<uint32_t>(wg * HC);
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    const uint32_t acc1_empty_ld = mapa_shared(smem_u32(&acc1_empty[wg]), 0);
    const uint32_t acc2_empty_ld = mapa_shared(smem_u32(&acc2_empty), 0);
    uint32_t nuse = 0;
    int g0 = 0;                                              // running chunk index of chunk 0 of the current super-tile
    uint32_t t_par = 0;
    long long a_a1 = 0, a_hin = 0, a_a2 = 0, a_math = 0, a_out = 0;
    const long long a_t0 = clock64();
    const bool acct = acct_cta && q == 0 && lane == 0;
    for (int pr = pair_first; pr < p.pairs; pr += pair_stride, t_par ^= 1u) {
      const int64_t grow = static_cast<int64_t>(pr * 2 + static_cast<int>(rank)) * BM + r;
      const bool row_ok = grow < p.M;
      for (int j = wg; j < NC; j += 2, ++nuse) {
        const uint32_t par = nuse & 1u;
        const int g = g0 + j;
        const int hbi = g % NHB;
        const uint32_t hpar = static_cast<uint32_t>(g / NHB) & 1u;
        const uint32_t hrow = hb(hbi) + r * 128;
        PW(smem_u32(&acc1_full[wg]), par, a_a1);
        PW(smem_u32(&hin_full[hbi]), hpar, a_hin);           // hpre chunk landed in the operand buffer
        tc_fence_after();
        const long long tm0 = kAcct ? clock64() : 0;
        uint32_t v[2][16];
        uint4 iv[2][2];
        const uint32_t swz_blk = hrow | (sw << 4);
        auto issue = [&](int s, int h4, uint32_t (&vv)[16], uint4 (&ii)[2]) {
          tmem_ld_32x32b_x16(t_acc1 + s * 16, vv);
          const uint32_t blk = swz_blk + (s >> 2) * KB_BYTES;
          ii[0] = ld_shared_v4(blk ^ (32u * h4));
          ii[1] = ld_shared_v4(blk ^ (32u * h4 + 16u));
        };
        auto process = [&](int s, int h4, const uint32_t (&vv)[16], const uint4 (&ii)[2]) {
          const uint32_t blk = swz_blk + (s >> 2) * KB_BYTES;
          const uint32_t adr_a = blk ^ (32u * h4), adr_b = blk ^ (32u * h4 + 16u);
          const uint32_t iw[8] = {ii[0].x, ii[0].y, ii[0].z, ii[0].w, ii[1].x, ii[1].y, ii[1].z, ii[1].w};
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = pack2_bf16(mul2(pk2u(vv[2 * i], vv[2 * i + 1]), gelu_bwd2(unpack2_bf16(iw[i]))));
          st_shared_v4(adr_a, o[0], o[1], o[2], o[3]);
          st_shared_v4(adr_b, o[4], o[5], o[6], o[7]);
        };
        issue(0, 0, v[0], iv[0]);
#pragma unroll 1
        for (int s = 0; s < HC / 16; s += 4) {
          tmem_ld_wait();
          issue(s + 1, 1, v[1], iv[1]);
          process(s, 0, v[0], iv[0]);
          tmem_ld_wait();
          issue(s + 2, 2, v[0], iv[0]);
          process(s + 1, 1, v[1], iv[1]);
          tmem_ld_wait();
          issue(s + 3, 3, v[1], iv[1]);
          process(s + 2, 2, v[0], iv[0]);
          tmem_ld_wait();
          if (s + 4 < HC / 16) issue(s + 4, 0, v[0], iv[0]);
          process(s + 3, 3, v[1], iv[1]);
        }
        tc_fence_before();                                   // this thread's TMEM reads are complete (tcgen05.wait::ld above)
        fence_proxy_async_smem();                            // its operand-buffer writes -> visible to UMMA / TMA
        __syncwarp();                                        // ... for all 32 rows of the warp
        if (lane == 0) {
          mbar_arrive_cluster(acc1_empty_ld);                // accumulator drained (leader's barrier)
          mbar_arrive(smem_u32(&h_full[hbi]));               // local: the storer may save the chunk
          mbar_arrive_cluster(mapa_shared(smem_u32(&h_ready[hbi]), 0));   // leader: GEMM2 may read it
        }
        if (kAcct) a_math += clock64() - tm0;
      }
      {
        // ---- output tile: acc2 -> bf16 dx, 32 bytes per row per step straight to global; each warpgroup takes 128 columns ----
        PW(smem_u32(&acc2_full), t_par, a_a2);
        tc_fence_after();
        const long long to0 = kAcct ? clock64() : 0;
        bf16* out_row = p.out + grow * DM;
        uint32_t v[2][16];
        auto process = [&](int s, const uint32_t (&vv)[16]) {
          if (row_ok) {
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = pack_bf16(__uint_as_float(vv[2 * i]), __uint_as_float(vv[2 * i + 1]));
            st_global_v4(out_row + s * 16, o[0], o[1], o[2], o[3]);
            st_global_v4(out_row + s * 16 + 8, o[4], o[5], o[6], o[7]);
          }
        };
        const int s_lo = wg * (DM / 32), s_hi = s_lo + DM / 32;
        tmem_ld_32x32b_x16(t_lane + ACC2_COL + s_lo * 16, v[0]);
#pragma unroll 1
        for (int s = s_lo; s < s_hi; s += 2) {
          tmem_ld_wait();
          tmem_ld_32x32b_x16(t_lane + ACC2_COL + (s + 1) * 16, v[1]);
          process(s, v[0]);
          tmem_ld_wait();
          if (s + 2 < s_hi) tmem_ld_32x32b_x16(t_lane + ACC2_COL + (s + 2) * 16, v[0]);
          process(s + 1, v[1]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc2_empty_ld);
        if (kAcct) a_out += clock64() - to0;
      }
      g0 += NC;
    }
    if (acct)
      printf("mlp pair acct: epilogue wg %d total %lld clk, waiting acc1_full %lld, hin_full %lld, acc2_full %lld; chunk math %lld, output tile %lld\n",
             wg, clock64() - a_t0, a_a1, a_hin, a_a2, a_math, a_out);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                       // nobody tears down while the peer may still signal / read it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
  }
}

}  // namespace

// sc_mlp_bwd_dgrad's pair path (mlp_fused.cu dispatches here): returns SC_OK or an error code
int sc_mlp_bwd_pair_launch(const void* dy, const void* w1, const void* w2, const void* hpre, void* dhpre, void* dx, int64_t M,
                           int R, int dbg, cudaStream_t st) {
  int rc;
  CUtensorMap tmX, tmWa, tmWb, tmH, tmHin;
  if ((rc = sc_encode_tmap_2d_bf16(&tmX, dy, DM, (uint64_t)M, DM * 2, 64, BM)) != SC_OK) return rc;
  // GEMM1: dh = dy W2, B(n = hidden, k = d) is W2 [256, R] read MN-major (inner = hidden, outer = d)
  if ((rc = sc_encode_tmap_2d_bf16(&tmWa, w2, (uint64_t)R, DM, (uint64_t)R * 2, 64, 64)) != SC_OK) return rc;
  // GEMM2: dx = dhpre W1, B(n = d, k = hidden) is W1 [R, 256] read MN-major (inner = d, outer = hidden)
  if ((rc = sc_encode_tmap_2d_bf16(&tmWb, w1, DM, (uint64_t)R, DM * 2, 64, 64)) != SC_OK) return rc;
  if ((rc = sc_encode_tmap_2d_bf16(&tmH, dhpre, (uint64_t)R, (uint64_t)M, (uint64_t)R * 2, 64, BM)) != SC_OK) return rc;
  if ((rc = sc_encode_tmap_2d_bf16(&tmHin, hpre, (uint64_t)R, (uint64_t)M, (uint64_t)R * 2, 64, BM)) != SC_OK) return rc;
  PairParams p;
  p.M = M; p.R = R; p.out = static_cast<bf16*>(dx); p.dbg = dbg;
  p.pairs = static_cast<int>((M + 2 * BM - 1) / (2 * BM));
  static bool attr_set_dev[64] = {};
  const int dev = sc_device_index();
  if (dev < 0) { sc_set_error("sc_mlp_bwd_dgrad: cudaGetDevice failed"); return SC_E_LAUNCH; }
  if (!attr_set_dev[dev]) {
    cudaError_t e = cudaFuncSetAttribute(mlp_bwd_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
    if (e != cudaSuccess) {
      sc_set_error("sc_mlp_bwd_dgrad: cannot set %d B dynamic smem: %s", SMEM_TOTAL, cudaGetErrorString(e));
      return SC_E_LAUNCH;
    }
    attr_set_dev[dev] = true;
  }
  int grid = sc_num_sms() & ~1;
  if (2 * p.pairs < grid) grid = 2 * p.pairs;
  cudaError_t e = sc_launch_kernel_cluster(mlp_bwd_pair_kernel, 2, dim3(grid), dim3(THREADS), SMEM_TOTAL, st, tmX, tmWa, tmWb, tmH,
                                           tmHin, p);
  sc_count_launch(1);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    sc_set_error("sc_mlp_bwd_dgrad (pair): launch failed: %s", cudaGetErrorString(e));
    return SC_E_LAUNCH;
  }
  return SC_OK;
}
