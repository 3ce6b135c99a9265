// common.cuh — shared device/host helpers for libsemcom_sm100.so (sm_100a only).
//
// PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit /
// ld / fences) and small math helpers.  Everything here is hand-written inline PTX; no
// CUTLASS/CuTe types are used.  Descriptor bit layouts follow the PTX ISA "tcgen05
// matrix descriptor" / "instruction descriptor" tables.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/semcom_sm100.h"

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

// ----------------------------------------------------------------------------------
// host-side error plumbing (api.cu owns the storage)
// ----------------------------------------------------------------------------------
void sc_set_error(const char* fmt, ...);
int sc_check_device();                 // SC_OK or SC_E_ARCH (device is not sm_100)
int sc_num_sms();
int sc_device_index();               // current device (0..63), -1 on error; function attributes are per device
void sc_count_launch(int n);           // feeds sc_launch_count()
int sc_encode_tmap_2d_bf16(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer,
                           uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer);

#define SC_CHECK_LAUNCH(name)                                                        \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      sc_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));          \
      return SC_E_LAUNCH;                                                            \
    }                                                                                \
  } while (0)

#define SC_REQUIRE(cond, code, ...)                                                  \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      sc_set_error(__VA_ARGS__);                                                     \
      return code;                                                                   \
    }                                                                                \
  } while (0)

static inline bool sc_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ----------------------------------------------------------------------------------
// launches: every kernel of the library goes through sc_launch_kernel.  With programmatic dependent launch (PDL,
// opt-in: sc_debug_set("pdl", 1) or SC_PDL=1; measured slightly slower on the PPO step, so off by default) a kernel may become resident while its predecessor in
// the stream is still draining: its prologue (barrier init, TMEM allocation, descriptor prefetch) overlaps the
// predecessor's tail and the launch latency disappears.  The contract every kernel keeps: NO global-memory access
// before pdl_wait(), and every CTA executes pdl_wait() -> completion of a grid implies completion of all earlier ones.
// ----------------------------------------------------------------------------------
int sc_pdl_enabled();                  // api.cu
int sc_take_direction();               // api.cu: 0 = ascending rows, 1 = descending (sc_debug_set("alternate", 1))
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t sc_launch_kernel_cluster(void (*kern)(KArgs...), int cluster_x, dim3 grid, dim3 block,
                                                   size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster_x;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (sc_pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t sc_launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                           cudaStream_t st, Args&&... args) {
  return sc_launch_kernel_cluster(kern, 1, grid, block, smem, st, static_cast<Args&&>(args)...);
}
#endif

// ----------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// ---- programmatic dependent launch ------------------------------------------------------
// pdl_wait: blocks until the preceding grid in the stream has completed and its memory is visible (no-op when the
// kernel was launched without the PDL attribute).  pdl_trigger: lets the NEXT kernel in the stream become resident.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_wait(); pdl_trigger(); }

// ---- mbarrier ---------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait with a suspend-time hint (ns): the thread sleeps in hardware until the phase completes
// or the hint expires, instead of hot-spinning and stealing issue slots from the epilogue warps.
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity), "r"(20000u)
      : "memory");
  return done;
}
// Bounded wait: a protocol bug must trap (-> CUDA error on the host), never hang the GPU box.
// SC_MBAR_SPIN_LIMIT (build.py: SC_MBAR_LIMIT=<log2>) shortens the bound for bring-up runs of new barrier protocols
// (every try may sleep up to 20 us, so 2^24 tries can be minutes).
#ifndef SC_MBAR_SPIN_LIMIT
#define SC_MBAR_SPIN_LIMIT (1u << 24)
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it) {
    if (it > SC_MBAR_SPIN_LIMIT) {                           // >= seconds even if every try returns at once
      printf("semcom: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// named barrier among `nthreads` threads of the CTA (id 1..15; id 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Non-blocking phase test.  Issued one stage AHEAD by the single-thread roles: the SYNCS round trip (~150 cycles even
// for a completed phase) then overlaps the MMA / TMA issue of the current stage instead of preceding it.
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}

// Hot-spinning wait (mbarrier.test_wait never suspends the thread).  Only for roles that are ONE thread per CTA
// (TMA producer, MMA issuer): the spin costs a few issue slots on one scheduler but reacts within tens of cycles.
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t it = 0; !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (it > (1u << 28)) {
      printf("semcom: mbarrier timeout (spin wait, block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// One lane of a CONVERGED warp (PTX elect.sync).  The single-thread roles (TMA producer, MMA issuer) run their loops
// with all 32 lanes converged and guard only the issuing instruction with this predicate: ptxas then keeps the
// warp-uniform descriptors / coordinates in uniform registers.  A loop inside `if (lane == 0)` is divergent code to
// ptxas, which wraps every UTCHMMA / UTMALDG in an R2UR + ELECT waterfall (~48 instructions per MMA: the issuing
// thread, not the tensor pipe, then sets the pace).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0u;
}

// ---- proxies / fences ---------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- TMA ------------------------------------------------------------------------------
// L2 prefetch of one TMA box: the later cp.async.bulk.tensor of the same box hits L2, so the HBM latency is
// taken off the shared-memory ring (bytes in flight are no longer bounded by the ring depth).
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* tm, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* tm, uint32_t bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// Multicast load: the box lands at the same CTA-relative offset in every CTA of `cta_mask` and completes bytes on the
// mbarrier at the same offset in each of them (every destination CTA has posted its own expect_tx).
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst_smem, const CUtensorMap* tm, uint32_t bar, int32_t c0,
                                               int32_t c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src_smem, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// address of 16-byte piece `pc` (0..7) of row r in a [rows x 128 B] SWIZZLE_128B tile (1 KB-aligned base)
__device__ __forceinline__ uint32_t sw128(uint32_t base, int r, int pc) {
  return base + r * 128 + ((pc ^ (r & 7)) << 4);
}

// ---- 2-CTA (cta_group::2) helpers: CTA pair = cluster of 2, rank 0 is the MMA leader ---------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
// TMA load issued by either CTA of the pair; the bytes are accounted on `cluster_bar_addr` (the leader's barrier)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst_smem, const CUtensorMap* tm, uint32_t cluster_bar_addr,
                                                int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tm)), "r"(cluster_bar_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same smem offset in BOTH CTAs of the pair once all prior MMAs retire
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(static_cast<uint16_t>(3))
      : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem) {   // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// ---- TMEM alloc -----------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {      // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {       // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}

// ---- UMMA descriptors -------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_128B, version 1 (sm_100):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// instruction descriptor for kind::f16, bf16 x bf16 -> fp32
//   [4,6) c_fmt=1(F32) | [7,10) a_fmt=1(BF16) | [10,13) b_fmt=1 | 15 a_major | 16 b_major |
//   [17,23) N>>3 | [24,29) M>>4        (major: 0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t m, uint32_t n, uint32_t a_mn,
                                                       uint32_t b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) |
         ((m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One k-block for the single-CTA MMA issuer in ONE asm block: the non-blocking test of the NEXT stage's full barrier
// is issued first, then the four K=16 MMAs of this stage and the commit that frees its shared-memory slot; the
// predicate is only turned into a register at the end, so the ~150-cycle SYNCS round trip overlaps the MMA issue
// instead of stalling the loop.  Returns 1 when the next stage has already landed.
#define SC_UMMA4_BODY                                                         \
  "setp.ne.b32 p, %7, 0;\n\t"                                                  \
  "setp.eq.b32 t, 0, 0;\n\t"                                                   \
  "cvt.u64.u32 sa, %4;\n\t"                                                    \
  "cvt.u64.u32 sb, %5;\n\t"                                                    \
  "tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %6, p;\n\t"                \
  "add.u64 a, %2, sa;\n\t"                                                     \
  "add.u64 b, %3, sb;\n\t"                                                     \
  "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %6, t;\n\t"                  \
  "add.u64 a, a, sa;\n\t"                                                      \
  "add.u64 b, b, sb;\n\t"                                                      \
  "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %6, t;\n\t"                  \
  "add.u64 a, a, sa;\n\t"                                                      \
  "add.u64 b, b, sb;\n\t"                                                      \
  "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %6, t;\n\t"
#define SC_UMMA4_OPERANDS                                                                                       \
  "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(a_step), "r"(b_step), "r"(idesc), "r"(accumulate_first),          \
      "r"(commit_bar), "r"(next_bar), "r"(next_parity)
// kTest: test `next_bar` ahead of the MMAs (result returned); kCommit: commit to `commit_bar` after them.  Stages of
// more than four MMAs chain a <true,false> block with a <false,true> block.
// kMc: the commit arrives on `commit_bar` in BOTH CTAs of a 2-CTA cluster (operand stages filled by multicast TMA are
// free only when both consumers have read them).
#define SC_COMMIT_LOCAL "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
#define SC_COMMIT_MC2                                                                                            \
  "{\n\t.reg .b16 cm;\n\tmov.b16 cm, 3;\n\t"                                                                    \
  "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%8], cm;\n\t}\n\t"
template <bool kTest = true, bool kCommit = true, bool kMc = false>
__device__ __forceinline__ uint32_t umma_kblock4_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t a_step,
                                                         uint32_t b_step, uint32_t idesc, uint32_t accumulate_first,
                                                         uint32_t commit_bar, uint32_t next_bar, uint32_t next_parity) {
  uint32_t ready = 0;
  if (kTest && kCommit && kMc) {
    asm volatile("{\n\t.reg .pred p, q, t;\n\t.reg .b64 a, b, sa, sb;\n\t"
                 "mbarrier.test_wait.parity.shared::cta.b64 q, [%9], %10;\n\t" SC_UMMA4_BODY SC_COMMIT_MC2
                 "selp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(ready) : SC_UMMA4_OPERANDS : "memory");
  } else if (kTest && kCommit) {
    asm volatile("{\n\t.reg .pred p, q, t;\n\t.reg .b64 a, b, sa, sb;\n\t"
                 "mbarrier.test_wait.parity.shared::cta.b64 q, [%9], %10;\n\t" SC_UMMA4_BODY SC_COMMIT_LOCAL
                 "selp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(ready) : SC_UMMA4_OPERANDS : "memory");
  } else if (kTest) {
    asm volatile("{\n\t.reg .pred p, q, t;\n\t.reg .b64 a, b, sa, sb;\n\t"
                 "mbarrier.test_wait.parity.shared::cta.b64 q, [%9], %10;\n\t" SC_UMMA4_BODY
                 "selp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(ready) : SC_UMMA4_OPERANDS : "memory");
  } else if (kCommit && kMc) {
    asm volatile("{\n\t.reg .pred p, t;\n\t.reg .b64 a, b, sa, sb;\n\t" SC_UMMA4_BODY SC_COMMIT_MC2
                 "mov.u32 %0, 0;\n\t}"
                 : "=r"(ready) : SC_UMMA4_OPERANDS : "memory");
  } else if (kCommit) {
    asm volatile("{\n\t.reg .pred p, t;\n\t.reg .b64 a, b, sa, sb;\n\t" SC_UMMA4_BODY SC_COMMIT_LOCAL
                 "mov.u32 %0, 0;\n\t}"
                 : "=r"(ready) : SC_UMMA4_OPERANDS : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p, t;\n\t.reg .b64 a, b, sa, sb;\n\t" SC_UMMA4_BODY
                 "mov.u32 %0, 0;\n\t}"
                 : "=r"(ready) : SC_UMMA4_OPERANDS : "memory");
  }
  return ready;
}
#undef SC_UMMA4_BODY
#undef SC_UMMA4_OPERANDS
#undef SC_COMMIT_LOCAL
#undef SC_COMMIT_MC2
// the same, arriving on the barrier at this offset in BOTH CTAs of a 2-CTA cluster (cta_group::1 MMAs)
__device__ __forceinline__ void umma_commit_mc2(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(static_cast<uint16_t>(3))
      : "memory");
}
// all prior tcgen05.mma of this thread arrive on `bar` when complete (implies fence::before)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

// ---- TMEM -> registers ----------------------------------------------------------------
// 32 lanes x 32-bit, 32 consecutive columns: thread t of the warp gets lane (base_lane + t),
// v[j] = column (col + j).  The warp may only touch lanes 32*(warp_id%4) .. +31.
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- vector global access -----------------------------------------------------------------
__device__ __forceinline__ void red_add_v4_f32(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  bf162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  bf162 t = *reinterpret_cast<bf162*>(&u);
  return __bfloat1622float2(t);
}

// ---- math ------------------------------------------------------------------------------------
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// GELU, tanh form (SPEC v0): 0.5 x (1 + tanh(c0 (x + c1 x^3)))
#define SC_GELU_C0 0.7978845608028654f
#define SC_GELU_C1 0.044715f
template <bool kFast>
__device__ __forceinline__ float gelu_fwd(float x) {
  float u = SC_GELU_C0 * x * fmaf(SC_GELU_C1, x * x, 1.0f);
  float t = kFast ? tanh_fast(u) : tanhf(u);
  return 0.5f * x * (1.0f + t);
}
template <bool kFast>
__device__ __forceinline__ float gelu_bwd(float x) {
  float x2 = x * x;
  float u = SC_GELU_C0 * x * fmaf(SC_GELU_C1, x2, 1.0f);
  float t = kFast ? tanh_fast(u) : tanhf(u);
  float du = SC_GELU_C0 * fmaf(3.0f * SC_GELU_C1, x2, 1.0f);
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
}

// one element of torch.optim.Adam (arithmetic order of the foreach implementation, lerp form for the first moment);
// shared by sc_adam_flat and the peer-memory reduce + Adam kernel so both paths round identically
__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float step_size, float b1, float b2,
                                            float eps, float inv_sqrt_bc2) {
  m = m + (1.f - b1) * (g - m);
  v = b2 * v + (1.f - b2) * g * g;
  const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
  p = p - step_size * (m / denom);
}
// bias corrections of step t, recomputed in double like the host does (device-side step counters: CUDA-graph replays)
__device__ __forceinline__ void adam_bias_corrections(double t, float lr, float b1, float b2, float& step_size,
                                                      float& inv_sqrt_bc2) {
  step_size = static_cast<float>(static_cast<double>(lr) / (1.0 - pow(static_cast<double>(b1), t)));
  inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(1.0 - pow(static_cast<double>(b2), t)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

#endif  // __CUDACC__
