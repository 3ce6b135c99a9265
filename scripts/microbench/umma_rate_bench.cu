// Raw tcgen05.mma issue-rate microbenchmark: one CTA per SM, one thread issues back-to-back
// 128 x N x 16 bf16 MMAs (operands: stale shared memory, SWIZZLE_128B K-major descriptors), commits once and waits.
// Prints clocks per MMA and the implied chip TFLOP/s at the measured SM clock.
#include "../../multimodal_sc_b200/csrc/common.cuh"

template <int N, int STRIDE_STAGES>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int n_mma) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t base = smem_u32(smem);
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 32) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const int st = (i >> 2) % STRIDE_STAGES;          // walk over a 4-stage ring like the real kernel
      const uint32_t a = base + st * 49152 + (i & 3) * 32;
      const uint32_t b = a + 16384;
      umma_bf16_ss(tmem + ((i >> 6) & 1) * 256, umma_smem_desc(a, 16, 1024), umma_smem_desc(b, 16, 1024), idesc, (i & 63) ? 1u : 0u);
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

template <int N, int SS>
void run(const char* name) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  auto k = bench<N, SS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 49152);
  const int n = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<148, 128, 4 * 49152>>>(d, n);
  cudaEventRecord(e0);
  k<<<148, 128, 4 * 49152>>>(d, n);
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
  run<256, 4>("ring of 4 stages");
  run<256, 1>("one stage");
  run<128, 4>("ring of 4 stages");
  run<64, 4>("ring of 4 stages");
  return 0;
}
