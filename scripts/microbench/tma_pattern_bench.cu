// TMA read-pattern microbenchmark: what HBM read bandwidth does a persistent one-CTA-per-SM kernel reach when it streams a
// row-major [M, 256] bf16 matrix (512-byte rows) into shared memory
//   (a) as the GEMM A operand does: 128-row x 64-column SWIZZLE_128B boxes (a quarter of every row per request, the four
//       k-blocks of a 128-row tile requested one after the other),
//   (b) as 1-D bulk copies of the same bytes (a 128-row tile of a [M, 256] matrix is one contiguous 64 KB block),
//   (c) as (a) but with the four k-blocks of a tile fetched by ONE 3-D request (64 columns x 4 chunks x 128 rows), which
//       walks every row's 512 bytes back to back?
// and how do the answers depend on the ring depth (stages in flight) and on pacing (a delay between requests that mimics
// the MMA consumption rate)?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_pattern_bench tma_pattern_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int KB = 16384;       // one k-block: 128 rows x 64 bf16

// mode 0: 2-D boxes, mode 1: 1-D bulk, mode 2: one 3-D request per tile (needs stages % 4 == 0)
__global__ void __launch_bounds__(128, 1)
stream_kernel(const __grid_constant__ CUtensorMap tm2, const __grid_constant__ CUtensorMap tm3, const uint8_t* base, int tiles,
              int stages, int mode, int delay, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[16];
  const uint32_t sb = smem_u32(smem);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 16; ++s) mbar_init(smem_u32(&full[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  // single thread: issue up to `stages` requests ahead, consume in order (consumption = waiting for the bytes).
  // 32-bit incremental bookkeeping only: the issue loop must not be what is measured
  const int per_req = (mode == 2) ? 4 : 1;                    // k-blocks per request
  const int slots = stages / per_req;
  const int rpt = 4 / per_req;                                // requests per tile
  const int my_tiles = (tiles - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1;
  const int total = my_tiles * rpt;
  int issued = 0, done = 0, islot = 0, dslot = 0, itile = static_cast<int>(blockIdx.x), ikb = 0;
  uint32_t dphase = 0;
  unsigned long long acc = 0;
  while (done < total) {
    while (issued < total && issued - done < slots) {
      const uint32_t bar = smem_u32(&full[islot]);
      const uint32_t dst = sb + islot * per_req * KB;
      mbar_expect_tx(bar, per_req * KB);
      if (mode == 0) tma_2d(dst, &tm2, bar, ikb * 64, itile * 128);
      else if (mode == 1) bulk_1d(dst, base + (static_cast<size_t>(itile) * 4 + ikb) * KB, KB, bar);
      else tma_3d(dst, &tm3, bar, 0, 0, itile * 128);
      ++issued;
      if (++islot == slots) islot = 0;
      if (++ikb == rpt) { ikb = 0; itile += static_cast<int>(gridDim.x); }
      if (delay > 0) { const long long t0 = clock64(); while (clock64() - t0 < delay) {} }
    }
    mbar_wait(smem_u32(&full[dslot]), dphase);
    acc += *reinterpret_cast<volatile uint32_t*>(smem + dslot * per_req * KB);
    ++done;
    if (++dslot == slots) { dslot = 0; dphase ^= 1u; }
  }
  if (acc == 0x12345678abcdefULL) *sink = acc;
}

// ---- store side: W warps per CTA, each owns two staging slots and streams (rows x 64-column) SWIZZLE_128B boxes of a row-major
// [M, ncols] bf16 matrix with TMA stores (the GEMM's staged epilogue writes 32-row x 64-column = 4 KB boxes, one per warp
// and chunk).  rows_per_box = 32 (4 KB) or 128 (16 KB).  No math, no loads: the write path alone.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__global__ void __launch_bounds__(256, 1)
store_kernel(const __grid_constant__ CUtensorMap tmS, int row_tiles, int col_chunks, int rows_per_box, int warps, int delay) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= warps || lane != 0) return;
  const uint32_t slot0 = smem_u32(smem) + warp * 2 * rows_per_box * 128;
  const int boxes_per_tile = 128 / rows_per_box;              // row boxes per 128-row tile
  int n = 0;
  for (int tile = blockIdx.x; tile < row_tiles; tile += gridDim.x) {
    // this warp's share of the tile: row boxes warp, warp + warps, ... ; every 64-column chunk
    for (int rb = warp; rb < boxes_per_tile; rb += warps)
      for (int c = 0; c < col_chunks; ++c, ++n) {
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");      // the slot used two stores ago has been read
        tma_store_2d(&tmS, slot0 + (n & 1) * rows_per_box * 128, c * 64, tile * 128 + rb * rows_per_box);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (delay > 0) { const long long t0 = clock64(); while (clock64() - t0 < delay) {} }
      }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const long long M = 1 << 21;                                  // 2 Mi rows x 512 B = 1 GiB
  uint8_t* buf;
  CK(cudaMalloc(&buf, M * 512));
  CK(cudaMemset(buf, 1, M * 512));
  uint8_t* flush;
  CK(cudaMalloc(&flush, 512ll << 20));
  unsigned long long* sink;
  CK(cudaMalloc(&sink, 8));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeFn enc = reinterpret_cast<EncodeFn>(fn);
  CUtensorMap tm2, tm3;
  {
    cuuint64_t dims[2] = {256, (cuuint64_t)M}; cuuint64_t str[1] = {512}; cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode 2d failed %d\n", (int)r); return 1; }
  }
  {
    // (64 columns, 4 chunks of 64 columns, M rows): the chunk stride is 128 bytes, the row stride 512 bytes
    cuuint64_t dims[3] = {64, 4, (cuuint64_t)M}; cuuint64_t str[2] = {128, 512}; cuuint32_t box[3] = {64, 4, 128}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode 3d failed %d\n", (int)r); return 1; }
  }
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * KB));
  const int tiles = (int)(M / 128);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const char* names[3] = {"2-D boxes 128x64 (GEMM A operand)", "1-D bulk 16 KB", "3-D request, whole 128x256 tile"};
  printf("# %d SMs, 1 GiB [M, 256] bf16, one issuing thread per CTA; GB/s of HBM reads\n", sms);
  for (int mode = 0; mode < 3; ++mode)
    for (int stages : {3, 4, 8, 12})
      for (int delay : {0, 400, 800}) {
        if (mode == 2 && stages % 4) continue;
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
          CK(cudaMemsetAsync(flush, rep, 512ll << 20));
          CK(cudaEventRecord(e0));
          stream_kernel<<<sms, 128, 12 * KB>>>(tm2, tm3, buf, tiles, stages, mode, delay, sink);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          CK(cudaGetLastError());
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          if (rep > 0 && ms < best) best = ms;
        }
        printf("%-40s stages %2d delay %4d clk: %7.1f us  %7.1f GB/s\n", names[mode], stages, delay, best * 1e3, M * 512.0 / (best * 1e-3) / 1e9);
      }
  // ---- store side ----
  printf("# TMA stores into a row-major [M, ncols] bf16 matrix (1 GiB), GB/s of HBM writes\n");
  CK(cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  for (int ncols : {256, 1024})
    for (int rows_per_box : {32, 128}) {
      CUtensorMap tmS;
      const long long Ms = (1ll << 30) / (ncols * 2);
      cuuint64_t dims[2] = {(cuuint64_t)ncols, (cuuint64_t)Ms}; cuuint64_t str[1] = {(cuuint64_t)ncols * 2};
      cuuint32_t box[2] = {64, (cuuint32_t)rows_per_box}; cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&tmS, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode store map failed %d\n", (int)r); return 1; }
      for (int warps : {1, 4, 8}) {
        if (rows_per_box == 128 && warps > 1) continue;
        for (int delay : {0, 400}) {
          float best = 1e9f;
          for (int rep = 0; rep < 4; ++rep) {
            CK(cudaMemsetAsync(flush, rep, 512ll << 20));
            CK(cudaEventRecord(e0));
            store_kernel<<<sms, 256, 65536>>>(tmS, (int)(Ms / 128), ncols / 64, rows_per_box, warps, delay);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
          }
          printf("stores: ncols %4d, box %3d rows x 64 cols (%2d KB), %d issuing warps, delay %3d clk: %7.1f us  %7.1f GB/s\n", ncols,
                 rows_per_box, rows_per_box * 128 / 1024, warps, delay, best * 1e3, (double)(1ll << 30) / (best * 1e-3) / 1e9);
        }
      }
    }
  return 0;
}
