// Raw tcgen05.mma issue-rate microbenchmark: one CTA per SM, one thread issues back-to-back
// 128 x N x 16 bf16 MMAs (operands: stale shared memory, SWIZZLE_128B K-major descriptors), commits once and waits.
// Prints clocks per MMA and the implied chip TFLOP/s at the measured SM clock.
#include "../../multimodal_sc_b200/csrc/common.cuh"

template <int N, int STRIDE_STAGES>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int n_mma, int fill, int commit_every) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t dummy[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&dummy[i]), 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t base = smem_u32(smem);
  if (fill) {                                           // random bf16 operands in [-1, 1): realistic bit toggling
    uint32_t x = 0x9E3779B9u * (blockIdx.x * 128 + threadIdx.x + 1);
    for (int i = threadIdx.x; i < 4 * 49152 / 4; i += 128) {
      x ^= x << 13; x ^= x >> 17; x ^= x << 5;
      const uint32_t lo = 0x3F000000u | ((x & 0x7Fu) << 16) | ((x >> 7 & 1u) << 31);       // bf16 in the high half
      const uint32_t hi = 0x3F000000u | ((x >> 8 & 0x7Fu) << 16) | ((x >> 15 & 1u) << 31);
      reinterpret_cast<uint32_t*>(smem)[i] = (lo >> 16) | (hi & 0xFFFF0000u);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 32) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const int st = (i >> 2) % STRIDE_STAGES;          // walk over a 4-stage ring like the real kernel
      const uint32_t a = base + st * 49152 + (i & 3) * 32;
      const uint32_t b = a + 16384;
      umma_bf16_ss(tmem + ((i >> 6) & 1) * 256, umma_smem_desc(a, 16, 1024), umma_smem_desc(b, 16, 1024), idesc, (i & 63) ? 1u : 0u);
      if (commit_every && (i & (commit_every - 1)) == commit_every - 1) umma_commit(smem_u32(&dummy[(i >> 2) & 7]));
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

// latency probes: (a) 4 MMAs + commit, then wait for the commit's mbarrier before the next group (no overlap):
// iteration time - 4 x 128 clk = commit-to-observed latency.  (b) two warps ping-pong on a pair of mbarriers.
__global__ void __launch_bounds__(128, 1) latency(long long* out, int iters, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar, ping, pong;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&ping), 1); mbar_init(smem_u32(&pong), 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot, base = smem_u32(smem);
  if (mode == 0 && threadIdx.x == 32) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 256, 0, 0);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_ss(tmem, umma_smem_desc(base + k * 32, 16, 1024), umma_smem_desc(base + 16384 + k * 32, 16, 1024), idesc, 1u);
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), i & 1);
    }
    out[blockIdx.x] = clock64() - t0;
  }
  if (mode == 1) {
    if (threadIdx.x == 32) {
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        mbar_arrive(smem_u32(&ping));
        mbar_wait(smem_u32(&pong), i & 1);
      }
      out[blockIdx.x] = clock64() - t0;
    } else if (threadIdx.x == 64) {
      for (int i = 0; i < iters; ++i) {
        mbar_wait(smem_u32(&ping), i & 1);
        mbar_arrive(smem_u32(&pong));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

// The GEMM's two-role handshake without any data movement: a producer thread re-arms `full` as soon as `empty`
// flips, the MMA warp (converged, one elected lane) waits `full`, issues 4 MMAs, commits to `empty`.
// variant bits: 1 = no tcgen05.fence::after_thread_sync after the wait, 2 = the MMA warp is a single divergent lane,
//               4 = producer hot-spins (test_wait), 8 = MMA warp hot-spins
__device__ __forceinline__ bool leaderless_print(int variant, int lane, long long acc_mma) { return acc_mma > 0; }

template <int S>
__global__ void __launch_bounds__(128, 1) ring(long long* out, int n_kb, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[S], empty[S];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot, base = smem_u32(smem);
  if (warp == 0 && lane == 0) {
    int st = 0; uint32_t ph = 0;
    for (int kb = 0; kb < n_kb; ++kb) {
      if (variant & 4) mbar_wait_spin(smem_u32(&empty[st]), ph ^ 1u); else mbar_wait(smem_u32(&empty[st]), ph ^ 1u);
      mbar_arrive(smem_u32(&full[st]));
      if (++st == S) { st = 0; ph ^= 1u; }
    }
  } else if (warp == 1 && (!(variant & 2) || lane == 0)) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 256, 0, 0);
    int st = 0; uint32_t ph = 0;
    const long long t0 = clock64();
    long long acc_wait = 0, acc_mma = 0, acc_commit = 0, acc_rest = 0;
    for (int kb = 0; kb < n_kb; ++kb) {
      const long long c0 = clock64();
      if (variant & 16) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&full[st])), "r"(ph) : "memory");
        if (!done) mbar_wait(smem_u32(&full[st]), ph);
      } else if (variant & 8) mbar_wait_spin(smem_u32(&full[st]), ph); else mbar_wait(smem_u32(&full[st]), ph);
      if (!(variant & 1)) tc_fence_after();
      const long long c1 = clock64();
      const uint32_t a = base + (st % 4) * 49152;
      const bool leader = (variant & 2) ? true : elect_one_sync();
      long long c2 = c1, c3 = c1;
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem + ((kb >> 4) & 1) * 256, umma_smem_desc(a + k * 32, 16, 1024), umma_smem_desc(a + 16384 + k * 32, 16, 1024), idesc, (kb & 15) || k ? 1u : 0u);
        c2 = clock64();
        umma_commit(smem_u32(&empty[st]));
        c3 = clock64();
      }
      if (!(variant & 2)) __syncwarp();
      if (++st == S) { st = 0; ph ^= 1u; }
      const long long c4 = clock64();
      acc_wait += c1 - c0; acc_mma += c2 - c1; acc_commit += c3 - c2; acc_rest += c4 - c3;
    }
    if (lane == 0 || (variant & 2)) out[blockIdx.x] = clock64() - t0;
    if (blockIdx.x == 0 && leaderless_print(variant, lane, acc_mma))
      printf("    breakdown per k-block: wait+fence %lld, 4 MMA issue %lld, commit %lld, rest %lld\n", acc_wait / n_kb, acc_mma / n_kb, acc_commit / n_kb, acc_rest / n_kb);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int S>
void run_ring(int variant) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(ring<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 49152);
  const int n = 4096;
  ring<S><<<148, 128, 4 * 49152>>>(d, n, variant);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("ring handshake, %d stages, variant %2d: %.0f clk per k-block (512 = tensor pipe floor) (%s)\n", S, variant, (double)h[0] / n, cudaGetErrorString(e));
  cudaFree(d);
}

// how far can the issuing thread run ahead of the tensor pipe?  timestamps after each of 24 back-to-back MMAs
__global__ void __launch_bounds__(128, 1) issue_trace(long long* out, int with_commit) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t dbar[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&dbar[i]), 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot, base = smem_u32(smem);
  if (threadIdx.x == 32) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 256, 0, 0);
    long long t[25];
    t[0] = clock64();
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      umma_bf16_ss(tmem, umma_smem_desc(base + (i & 3) * 32, 16, 1024), umma_smem_desc(base + 16384 + (i & 3) * 32, 16, 1024), idesc, 1u);
      if (with_commit && (i & 3) == 3) umma_commit(smem_u32(&dbar[i >> 2]));
      t[i + 1] = clock64();
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long tend = clock64();
    if (blockIdx.x == 0) {
      for (int i = 0; i < 25; ++i) out[i] = t[i] - t[0];
      out[25] = tend - t[0];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

void run_issue_trace() {
  long long* d;
  cudaMalloc(&d, 32 * sizeof(long long));
  cudaFuncSetAttribute(issue_trace, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int wc = 0; wc < 2; ++wc) {
    issue_trace<<<1, 128, 65536>>>(d, wc);
    issue_trace<<<1, 128, 65536>>>(d, wc);
    cudaDeviceSynchronize();
    long long h[26];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("issue trace %s (clk after each of 24 MMAs of 128 clk):", wc ? "with a commit after every 4th MMA" : "no commits");
    for (int i = 1; i < 25; ++i) printf(" %lld", h[i]);
    printf(" | all complete at %lld\n", h[25]);
  }
  cudaFree(d);
}

void run_latency() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(latency, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int mode = 0; mode < 2; ++mode) {
    const int iters = 2000;
    latency<<<148, 128, 65536>>>(d, iters, mode);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    if (mode == 0) printf("4 MMAs + commit + wait: %.0f clk per iteration -> commit-to-observed latency %.0f clk (%s)\n", (double)h[0] / iters, (double)h[0] / iters - 512, cudaGetErrorString(e));
    else printf("mbarrier ping-pong between two warps: %.0f clk per round trip (%s)\n", (double)h[0] / iters, cudaGetErrorString(e));
  }
  cudaFree(d);
}

template <int N, int SS>
void run(const char* name, int n = 4096, int fill = 0, int commit_every = 0) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  auto k = bench<N, SS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 49152);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<148, 128, 4 * 49152>>>(d, n, fill, commit_every);
  cudaEventRecord(e0);
  k<<<148, 128, 4 * 49152>>>(d, n, fill, commit_every);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double clk = (double)h[0] / n;
  double fl = 148.0 * n * 2.0 * 128 * N * 16;
  printf("%-20s N=%3d: %.1f clk/MMA (SM0)  kernel %.3f ms -> %.0f TFLOP/s chip, implied SM clock %.0f MHz (%s)\n", name, N, clk, ms,
         fl / ms / 1e9, h[0] / ms / 1e3, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run_issue_trace();
  run_latency();
  for (int v : {0, 2, 16, 17, 18, 19}) run_ring<4>(v);
  run<256, 4>("ring of 4 stages");
  run<256, 1>("one stage");
  run<128, 4>("ring of 4 stages");
  run<64, 4>("ring of 4 stages");
  // sustained: ~100 ms of back-to-back MMAs -> the clock the tensor pipe holds under the power cap
  run<256, 4>("sustained 100 ms", 4096 * 300);
  run<256, 4>("sustained 400 ms", 4096 * 1200);
  run<256, 4>("random data 100 ms", 4096 * 300, 1);
  run<256, 4>("random data 400 ms", 4096 * 1200, 1);
  run<256, 4>("random data burst", 4096, 1);
  run<256, 4>("zeros, commit/4", 4096 * 8, 0, 4);
  run<256, 4>("zeros, commit/16", 4096 * 8, 0, 16);
  run<256, 4>("random, commit/4", 4096 * 8, 1, 4);
  run<256, 4>("random, commit/16", 4096 * 8, 1, 16);
  run<256, 4>("random, no commit", 4096 * 8, 1, 0);
  return 0;
}
