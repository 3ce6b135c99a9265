// dp_peer.cu — data-parallel gradient exchange fused with the optimizer over NVLink peer memory
// (include/semcom_sm100.h: sc_peer_*, sc_dp_barrier, sc_dp_reduce_adam).
//
// One process per GPU.  The flat gradient bucket, the fp32 master parameters and the bf16 shadow of every rank live in
// memory that every other rank has mapped (cudaIpc handles exchanged by the host once).  Instead of
//     ncclAllReduce(g32)  ->  adam(p32, g32, m, v) on every rank over the whole bucket
// rank r pulls slice r of every rank's gradients through NVLink (P2P loads), sums them in rank order, applies Adam with
// its shard of the moments and pushes the new parameters (fp32 + bf16) of the slice into every rank's buffers (P2P
// stores): reduce-scatter + sharded optimizer + all-gather in ONE pass over the bucket, framed by two flag barriers.
// The barriers are kernels of one CTA: thread p stores the new epoch into rank p's flag array (st.release.sys) and
// polls the local slot that rank p writes (ld.acquire.sys).  Stream order does the rest: everything a rank enqueued
// before its barrier has completed when the barrier kernel runs, and release / acquire at system scope are cumulative.
#include "common.cuh"

namespace {

constexpr int kDpThreads = 256;
constexpr int kDpUnroll = 2;                       // float4 per thread and pass, every rank's load issued before the first add
constexpr unsigned long long kBarrierTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

struct PeerPtrs {
  const float* g[SC_DP_MAX_WORLD];
  float* p32[SC_DP_MAX_WORLD];
  bf16* p16[SC_DP_MAX_WORLD];
};
struct FlagPtrs {
  unsigned long long* f[SC_DP_MAX_WORLD];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// peer loads: L2 only (a line of another GPU's memory must not be served from a stale L1 line of an earlier launch on
// this SM; L1 is invalidated between kernels, .cg makes it explicit)
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(32) dp_barrier_kernel(FlagPtrs fl, long long* epoch, int world, int rank) {
  pdl_sync();
  const unsigned long long e = static_cast<unsigned long long>(*epoch) + 1ull;
  const int p = threadIdx.x;
  if (p < world) {
    __threadfence_system();
    st_release_sys(fl.f[p] + rank, e);                           // "rank has arrived at barrier e", visible to rank p
    const unsigned long long* mine = fl.f[rank] + p;             // written by rank p
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(mine) < e) {
      __nanosleep(64);
      if (global_timer_ns() - t0 > kBarrierTimeoutNs) {
        printf("sc_dp_barrier: rank %d waited 20 s for rank %d at epoch %llu\n", rank, p, e);
        __trap();
      }
    }
    __threadfence_system();
  }
  __syncwarp();
  if (p == 0) *epoch = static_cast<long long>(e);
}

__global__ void __launch_bounds__(kDpThreads)
dp_reduce_adam_kernel(PeerPtrs pp, float* __restrict__ m, float* __restrict__ v, int world, int rank, int64_t lo4,
                      int64_t hi4, float step_size, float b1, float b2, float eps, float inv_sqrt_bc2, float lr,
                      int64_t step, const int64_t* __restrict__ step_dev) {
  pdl_sync();
  if (step_dev) {
    __shared__ float bc[2];
    if (threadIdx.x == 0) adam_bias_corrections(static_cast<double>(step + *step_dev), lr, b1, b2, bc[0], bc[1]);
    __syncthreads();
    step_size = bc[0];
    inv_sqrt_bc2 = bc[1];
  }
  const float* __restrict__ p_local = pp.p32[rank];
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t base = lo4 + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; base < hi4;
       base += stride * kDpUnroll) {
    float4 acc[kDpUnroll], par[kDpUnroll], mo[kDpUnroll], vo[kDpUnroll];
    // every load of the pass is issued before the first use: (world + 3) x kDpUnroll 16-byte requests in flight
#pragma unroll
    for (int u = 0; u < kDpUnroll; ++u) {
      const int64_t i = base + u * stride;
      if (i < hi4) {
        acc[u] = ld_peer_f4(pp.g[0] + 4 * i);
        par[u] = reinterpret_cast<const float4*>(p_local)[i];
        mo[u] = reinterpret_cast<const float4*>(m)[i];
        vo[u] = reinterpret_cast<const float4*>(v)[i];
      }
    }
    for (int r = 1; r < world; ++r) {
      float4 g[kDpUnroll];
#pragma unroll
      for (int u = 0; u < kDpUnroll; ++u) {
        const int64_t i = base + u * stride;
        if (i < hi4) g[u] = ld_peer_f4(pp.g[r] + 4 * i);
      }
#pragma unroll
      for (int u = 0; u < kDpUnroll; ++u) {
        const int64_t i = base + u * stride;
        if (i < hi4) { acc[u].x += g[u].x; acc[u].y += g[u].y; acc[u].z += g[u].z; acc[u].w += g[u].w; }
      }
    }
#pragma unroll
    for (int u = 0; u < kDpUnroll; ++u) {
      const int64_t i = base + u * stride;
      if (i >= hi4) continue;
      float* P = &par[u].x; float* G = &acc[u].x; float* Mo = &mo[u].x; float* V = &vo[u].x;
#pragma unroll
      for (int j = 0; j < 4; ++j) adam_update(P[j], G[j], Mo[j], V[j], step_size, b1, b2, eps, inv_sqrt_bc2);
      reinterpret_cast<float4*>(m)[i] = mo[u];
      reinterpret_cast<float4*>(v)[i] = vo[u];
      const uint2 sh = make_uint2(pack_bf16(par[u].x, par[u].y), pack_bf16(par[u].z, par[u].w));
      for (int r = 0; r < world; ++r) {
        reinterpret_cast<float4*>(pp.p32[r])[i] = par[u];
        reinterpret_cast<uint2*>(pp.p16[r])[i] = sh;
      }
    }
  }
}

inline void dp_slice(int64_t n4, int world, int rank, int64_t& lo, int64_t& hi) {
  const int64_t base = n4 / world, rem = n4 % world;
  lo = rank * base + (rank < rem ? rank : rem);
  hi = lo + base + (rank < rem ? 1 : 0);
}

}  // namespace

extern "C" {

int sc_peer_alloc(void** ptr, int64_t bytes) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(ptr && bytes > 0, SC_E_SHAPE, "sc_peer_alloc: bad arguments");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, static_cast<size_t>(bytes));
  SC_REQUIRE(e == cudaSuccess, SC_E_DRIVER, "sc_peer_alloc: cudaMalloc(%lld) failed: %s", (long long)bytes,
             cudaGetErrorString(e));
  e = cudaMemset(p, 0, static_cast<size_t>(bytes));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(p);
    sc_set_error("sc_peer_alloc: cudaMemset failed: %s", cudaGetErrorString(e));
    return SC_E_DRIVER;
  }
  *ptr = p;
  return SC_OK;
}

int sc_peer_free(void* ptr) {
  if (!ptr) return SC_OK;
  cudaError_t e = cudaFree(ptr);
  SC_REQUIRE(e == cudaSuccess, SC_E_DRIVER, "sc_peer_free: %s", cudaGetErrorString(e));
  return SC_OK;
}

int sc_peer_export(const void* ptr, void* handle) {
  static_assert(sizeof(cudaIpcMemHandle_t) == SC_PEER_HANDLE_BYTES, "handle size");
  SC_REQUIRE(ptr && handle, SC_E_SHAPE, "sc_peer_export: null pointer");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(ptr));
  SC_REQUIRE(e == cudaSuccess, SC_E_DRIVER, "sc_peer_export: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  memcpy(handle, &h, sizeof(h));
  return SC_OK;
}

int sc_peer_import(const void* handle, void** ptr) {
  SC_REQUIRE(ptr && handle, SC_E_SHAPE, "sc_peer_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  SC_REQUIRE(e == cudaSuccess, SC_E_DRIVER, "sc_peer_import: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
  *ptr = p;
  return SC_OK;
}

int sc_peer_close(void* ptr) {
  if (!ptr) return SC_OK;
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  SC_REQUIRE(e == cudaSuccess, SC_E_DRIVER, "sc_peer_close: %s", cudaGetErrorString(e));
  return SC_OK;
}

int sc_dp_barrier(void* const* flags, int64_t* epoch, int world, int rank, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(flags && epoch && world >= 1 && world <= SC_DP_MAX_WORLD && rank >= 0 && rank < world, SC_E_SHAPE,
             "sc_dp_barrier: world=%d rank=%d", world, rank);
  FlagPtrs fl;
  for (int r = 0; r < SC_DP_MAX_WORLD; ++r) fl.f[r] = nullptr;
  for (int r = 0; r < world; ++r) {
    SC_REQUIRE(flags[r] && (reinterpret_cast<uintptr_t>(flags[r]) & 7) == 0, SC_E_ALIGN, "sc_dp_barrier: flags[%d]", r);
    fl.f[r] = static_cast<unsigned long long*>(flags[r]);
  }
  sc_launch_kernel(dp_barrier_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, fl,
                   reinterpret_cast<long long*>(epoch), world, rank);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_dp_barrier");
  return SC_OK;
}

int sc_dp_reduce_adam(const void* const* g32, void* const* p32, void* const* p16, float* m, float* v, int world,
                      int rank, int64_t n, float lr, float beta1, float beta2, float eps, int64_t step,
                      const int64_t* step_dev, int max_ctas, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(g32 && p32 && p16 && m && v && world >= 1 && world <= SC_DP_MAX_WORLD && rank >= 0 && rank < world &&
                 n > 0 && n % 4 == 0 && (step >= 1 || (step_dev && step >= 0)),
             SC_E_SHAPE, "sc_dp_reduce_adam: world=%d rank=%d n=%lld step=%lld", world, rank, (long long)n,
             (long long)step);
  PeerPtrs pp;
  for (int r = 0; r < SC_DP_MAX_WORLD; ++r) { pp.g[r] = nullptr; pp.p32[r] = nullptr; pp.p16[r] = nullptr; }
  for (int r = 0; r < world; ++r) {
    SC_REQUIRE(g32[r] && p32[r] && p16[r], SC_E_SHAPE, "sc_dp_reduce_adam: null buffer of rank %d", r);
    SC_REQUIRE(sc_aligned16(g32[r]) && sc_aligned16(p32[r]) && (reinterpret_cast<uintptr_t>(p16[r]) & 7) == 0,
               SC_E_ALIGN, "sc_dp_reduce_adam: unaligned buffer of rank %d", r);
    pp.g[r] = static_cast<const float*>(g32[r]);
    pp.p32[r] = static_cast<float*>(p32[r]);
    pp.p16[r] = static_cast<bf16*>(p16[r]);
  }
  SC_REQUIRE(sc_aligned16(m) && sc_aligned16(v), SC_E_ALIGN, "sc_dp_reduce_adam: unaligned moments");
  int64_t lo4, hi4;
  dp_slice(n / 4, world, rank, lo4, hi4);
  const double tstep = (double)(step >= 1 ? step : 1);
  const float step_size = (float)((double)lr / (1.0 - pow((double)beta1, tstep)));
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)beta2, tstep)));
  int64_t blocks = (hi4 - lo4 + (int64_t)kDpThreads * kDpUnroll - 1) / ((int64_t)kDpThreads * kDpUnroll);
  const int64_t cap = max_ctas > 0 ? max_ctas : (int64_t)sc_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  sc_launch_kernel(dp_reduce_adam_kernel, dim3((unsigned)blocks), dim3(kDpThreads), 0, (cudaStream_t)stream, pp, m, v,
                   world, rank, lo4, hi4, step_size, beta1, beta2, eps, inv_sqrt_bc2, lr, step, step_dev);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_dp_reduce_adam");
  return SC_OK;
}

}  // extern "C"
