// elementwise.cu — memory-bound fused kernels: patchify+cast, LayerNorm fwd/bwd, column sums,
// mean-pool, fused MSE fwd+bwd, casts.  All 128-bit vectorised, coalesced, warp-shuffle
// reductions; grids are persistent (a multiple of the SM count) with grid-stride loops.
// SPEC v0 rows a1/a2 (patchify), a1/a2/a5 (LayerNorm), a5 (mean-pool), a6 (semantic distortion).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void load8(const bf16* p, float (&x)[8]) {
  uint4 w = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16(w.x), b = unpack_bf16(w.y), c = unpack_bf16(w.z), d = unpack_bf16(w.w);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
}
__device__ __forceinline__ void load8(const float* p, float (&x)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
__device__ __forceinline__ void store8(bf16* p, const float (&x)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]),
                                            pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
}
__device__ __forceinline__ void store8(float* p, const float (&x)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(x[4], x[5], x[6], x[7]);
}

inline int persistent_blocks(int64_t work_blocks, int per_sm) {
  int64_t cap = static_cast<int64_t>(sc_num_sms()) * per_sm;
  return static_cast<int>(work_blocks < cap ? (work_blocks > 0 ? work_blocks : 1) : cap);
}

template <typename T> struct Raw8;
template <> struct Raw8<bf16> { uint4 v; };
template <> struct Raw8<float> { float4 a, b; };

__device__ __forceinline__ void ld_raw(const bf16* p, Raw8<bf16>& r) { r.v = *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void ld_raw(const float* p, Raw8<float>& r) {
  r.a = *reinterpret_cast<const float4*>(p);
  r.b = *reinterpret_cast<const float4*>(p + 4);
}
__device__ __forceinline__ void cvt_raw(const Raw8<bf16>& r, float (&x)[8]) {
  float2 a = unpack_bf16(r.v.x), b = unpack_bf16(r.v.y), c = unpack_bf16(r.v.z), d = unpack_bf16(r.v.w);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
}
__device__ __forceinline__ void cvt_raw(const Raw8<float>& r, float (&x)[8]) {
  x[0] = r.a.x; x[1] = r.a.y; x[2] = r.a.z; x[3] = r.a.w; x[4] = r.b.x; x[5] = r.b.y; x[6] = r.b.z; x[7] = r.b.w;
}

// ------------------------------------------------------------------------------------------
// patchify + cast: one thread per 8 output features (always inside one image row)
// ------------------------------------------------------------------------------------------
template <typename TI>
__global__ void __launch_bounds__(kThreads)
patchify_kernel(const TI* __restrict__ img, bf16* __restrict__ out, int64_t groups, int C, int H,
                int W, int P, const int64_t* __restrict__ row_idx) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  const int Wt = W / P, Ht = H / P;
  const int K = C * P * P;
  const int K8 = K / 8;
  const int P8 = P / 8;
  // kUnroll independent 16-byte loads are issued before the first store: with one load in flight per thread the
  // kernel sat at 2.6 TB/s (latency-bound; every warp request touches 16 lines of a strided image row set)
  constexpr int kUnroll = 4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t g0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; g0 < groups; g0 += stride * kUnroll) {
    Raw8<TI> raw[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t g = g0 + u * stride;
      if (g < groups) {
        const int kg = static_cast<int>(g % K8);
        const int64_t tok = g / K8;
        const int px0 = (kg % P8) * 8;
        const int py = (kg / P8) % P;
        const int c = kg / (P8 * P);
        const int tx = static_cast<int>(tok % Wt);
        const int ty = static_cast<int>((tok / Wt) % Ht);
        int64_t b = tok / (Wt * Ht);
        if (row_idx) b = __ldg(row_idx + b);            // fused minibatch gather: output sample -> stored image
        ld_raw(img + ((b * C + c) * H + (ty * P + py)) * static_cast<int64_t>(W) + tx * P + px0, raw[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t g = g0 + u * stride;
      if (g < groups) {
        float x[8];
        cvt_raw(raw[u], x);
        store8(out + g * 8, x);
      }
    }
  }
}

// Compile-time geometry (the model's two observation shapes): every index is a shift / mask, one thread moves 16 bytes.
// The generic kernel above spends ~150 instructions of integer division per 16-byte load (ncu: issue-bound at
// 2.4 TB/s); here a CTA walks "token rows" (one image, one row of patches: C x P image rows of W pixels) in SOURCE order,
// so a warp reads whole 2W-byte image rows and scatters them as 2P-byte pieces of Wt neighbouring token rows.
template <typename TI, int C, int H, int W, int P>
__global__ void __launch_bounds__(kThreads)
patchify_fixed_kernel(const TI* __restrict__ img, bf16* __restrict__ out, int64_t token_rows,
                      const int64_t* __restrict__ row_idx) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  constexpr int Wt = W / P, Ht = H / P, K = C * P * P;
  constexpr int XG = W / 8;                     // 16-byte (8-pixel) groups per image row
  constexpr int GROUPS = C * P * XG;            // per token row
  constexpr int kUnroll = (GROUPS >= 4 * kThreads) ? 4 : ((GROUPS >= 2 * kThreads) ? 2 : 1);
  // small token rows (BEV: 64 groups): several token rows per CTA iteration so that every thread has a 16-byte piece
  constexpr int RPI = (GROUPS < kThreads) ? kThreads / GROUPS : 1;
  const int sub = (RPI > 1) ? static_cast<int>(threadIdx.x) / GROUPS : 0;
  const int tid = (RPI > 1) ? static_cast<int>(threadIdx.x) % GROUPS : static_cast<int>(threadIdx.x);
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * RPI + sub; r < token_rows; r += static_cast<int64_t>(gridDim.x) * RPI) {
    const int ty = static_cast<int>(r % Ht);
    int64_t b = r / Ht;
    const int64_t dst_row0 = r * Wt;            // first token of this token row (output order = (b, ty, tx))
    if (row_idx) b = __ldg(row_idx + b);
    const TI* src = img + (b * C * H + ty * P) * static_cast<int64_t>(W);
    constexpr int STEP = (RPI > 1) ? GROUPS : kThreads;
    for (int g0 = tid; g0 < GROUPS; g0 += STEP * kUnroll) {
      Raw8<TI> raw[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int g = g0 + u * STEP;
        if (g < GROUPS) {
          const int xg = g % XG, py = (g / XG) % P, c = g / (XG * P);
          ld_raw(src + (static_cast<int64_t>(c) * H + py) * W + xg * 8, raw[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int g = g0 + u * STEP;
        if (g < GROUPS) {
          const int xg = g % XG, py = (g / XG) % P, c = g / (XG * P);
          const int tx = (xg * 8) / P, pxin = (xg * 8) % P;
          float x[8];
          cvt_raw(raw[u], x);
          store8(out + (dst_row0 + tx) * K + (c * P + py) * P + pxin, x);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, lane owns NCH chunks of 8 contiguous elements (D = 256*NCH).
// Each warp handles R rows per iteration and issues all of their 16-byte loads before touching any
// of them: with 1 row in flight per warp the kernels sat at ~3 TB/s (latency-bound), R rows put
// R x more bytes in flight per SM.
// ------------------------------------------------------------------------------------------

template <typename T, int NCH, int R>
__global__ void __launch_bounds__(kThreads, 3)
layernorm_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, int64_t M, float eps) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  constexpr int D = 256 * NCH;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  float g[NCH][8], bta[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    load8(gamma + c * 256 + lane * 8, g[c]);
    load8(beta + c * 256 + lane * 8, bta[c]);
  }
  for (int64_t base = warp0 * R; base < M; base += nwarps * R) {
    Raw8<T> raw[R][NCH];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = min(base + r, M - 1);
#pragma unroll
      for (int c = 0; c < NCH; ++c) ld_raw(x + row * D + c * 256 + lane * 8, raw[r][c]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = base + r;
      if (row >= M) break;
      float v[NCH][8];
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        cvt_raw(raw[r][c], v[c]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[c][j];
      }
      const float mean = warp_sum(s) * (1.0f / D);
      float q = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[c][j] - mean;
          q = fmaf(d, d, q);
        }
      const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf((v[c][j] - mean) * rstd, g[c][j], bta[c][j]);
        store8(y + row * D + c * 256 + lane * 8, o);
      }
      if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
      }
    }
  }
}

// dx = dres + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
// dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy ; colsum += sum_rows dx
template <typename T, int NCH, int R>
__global__ void __launch_bounds__(kThreads, 2)
layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ gamma,
                     const T* __restrict__ dres, T* __restrict__ dx, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ colsum, int64_t M, int reverse) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  constexpr int D = 256 * NCH;
  __shared__ float red[3][D];
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) (&red[0][0])[i] = 0.f;
  __syncthreads();
  float gm[NCH][8];
  float a_dg[NCH][8], a_db[NCH][8], a_cs[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    load8(gamma + c * 256 + lane * 8, gm[c]);
#pragma unroll
    for (int j = 0; j < 8; ++j) a_dg[c][j] = a_db[c][j] = a_cs[c][j] = 0.f;
  }
  const int64_t last_base = (M - 1) / R * R;
  for (int64_t fbase = warp0 * R; fbase < M; fbase += nwarps * R) {
    const int64_t base = reverse ? last_base - fbase : fbase;       // descending row groups when asked to
    Raw8<T> rx[R][NCH], rdy[R][NCH], rres[R][NCH];
    float mu[R], rs[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = min(base + r, M - 1);
      mu[r] = mean[row];
      rs[r] = rstd[row];
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int64_t off = row * D + c * 256 + lane * 8;
        ld_raw(x + off, rx[r][c]);
        ld_raw(dy + off, rdy[r][c]);
        if (dres) ld_raw(dres + off, rres[r][c]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = base + r;
      if (row >= M) break;
      float xh[NCH][8], g[NCH][8];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        float xv[8], dyv[8];
        cvt_raw(rx[r][c], xv);
        cvt_raw(rdy[r][c], dyv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[c][j] = (xv[j] - mu[r]) * rs[r];
          g[c][j] = dyv[j] * gm[c][j];
          s1 += g[c][j];
          s2 = fmaf(g[c][j], xh[c][j], s2);
          a_dg[c][j] = fmaf(dyv[j], xh[c][j], a_dg[c][j]);
          a_db[c][j] += dyv[j];
        }
      }
      const float m1 = warp_sum(s1) * (1.0f / D);
      const float m2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rs[r] * (g[c][j] - m1 - xh[c][j] * m2);
        if (dres) {
          float rr[8];
          cvt_raw(rres[r][c], rr);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += rr[j];
        }
        store8(dx + row * D + c * 256 + lane * 8, o);
        if (colsum) {
#pragma unroll
          for (int j = 0; j < 8; ++j) a_cs[c][j] += o[j];
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = c * 256 + lane * 8 + j;
      atomicAdd(&red[0][col], a_dg[c][j]);
      atomicAdd(&red[1][col], a_db[c][j]);
      if (colsum) atomicAdd(&red[2][col], a_cs[c][j]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dgamma + i, red[0][i]);
    atomicAdd(dbeta + i, red[1][i]);
    if (colsum) atomicAdd(colsum + i, red[2][i]);
  }
}

// ------------------------------------------------------------------------------------------
// column sums of a bf16 [M, N] matrix (bias gradients)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
colsum_kernel(const bf16* __restrict__ x, float* __restrict__ out, int64_t M, int N_total, int64_t ld,
              int64_t rows_per_block) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  extern __shared__ float sred[];          // [<= 1024]
  const int col0 = blockIdx.y * 1024;
  const int N = min(1024, N_total - col0);  // columns handled by this block
  x += col0;
  out += col0;
  const int G = N / 8;                      // column groups
  const int rpi = blockDim.x / G;           // rows per iteration
  const int cg = threadIdx.x % G;
  const int rg = threadIdx.x / G;
  for (int i = threadIdx.x; i < N; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int64_t r0 = blockIdx.x * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  if (rg < rpi) {
    for (int64_t r = r0 + rg; r < r1; r += rpi) {
      float v[8];
      load8(x + r * ld + cg * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sred[cg * 8 + j], acc[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) atomicAdd(out + i, sred[i]);
}

// ------------------------------------------------------------------------------------------
// mean-pool over S tokens: one block of 4 warps per sample, D = 256 (lane owns 8 columns)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
meanpool_kernel(const bf16* __restrict__ x, bf16* __restrict__ out_bf16, float* __restrict__ out_f32,
                int64_t B, int S, int D) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  __shared__ float red[4][1024];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nch = D / 256;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    for (int c = 0; c < nch; ++c) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int s = warp; s < S; s += 4) {
        float v[8];
        load8(x + (b * S + s) * D + c * 256 + lane * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) red[warp][c * 256 + lane * 8 + j] = acc[j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      const float m = (red[0][i] + red[1][i] + red[2][i] + red[3][i]) * (1.0f / S);
      if (out_bf16) out_bf16[b * D + i] = __float2bfloat16(m);
      if (out_f32) out_f32[b * D + i] = m;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// fused MSE forward + backward
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
mse_kernel(const bf16* __restrict__ r, const bf16* __restrict__ t, const bf16* __restrict__ t1,
           float* __restrict__ loss, bf16* __restrict__ dr, int64_t groups, float inv_numel, float grad_scale,
           int seg_rows, int D8) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  __shared__ float wsum[kThreads / 32];
  float acc = 0.f;
  // four independent (recon, target) load pairs in flight per thread before the first use (latency-bound otherwise)
  constexpr int kUnroll = 4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t g0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; g0 < groups; g0 += stride * kUnroll) {
    Raw8<bf16> ra[kUnroll], rb[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t g = g0 + u * stride;
      if (g < groups) {
        ld_raw(r + g * 8, ra[u]);
        if (t1) {
          // r rows are (sample, modality, token); targets are two [samples*seg_rows, D] matrices
          const int64_t row = g / D8;
          const int c8 = static_cast<int>(g % D8);
          const int64_t smp = row / (2 * seg_rows);
          const int rem = static_cast<int>(row % (2 * seg_rows));
          const bf16* tb = rem < seg_rows ? t : t1;
          ld_raw(tb + ((smp * seg_rows + (rem % seg_rows)) * D8 + c8) * 8, rb[u]);
        } else {
          ld_raw(t + g * 8, rb[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t g = g0 + u * stride;
      if (g < groups) {
        float a[8], b[8], d[8];
        cvt_raw(ra[u], a);
        cvt_raw(rb[u], b);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float e = a[j] - b[j];
          acc = fmaf(e, e, acc);
          d[j] = 2.0f * e * grad_scale;
        }
        if (dr) store8(dr + g * 8, d);
      }
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kThreads / 32 ? wsum[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(loss, v * inv_numel);
  }
}

__global__ void __launch_bounds__(kThreads)
cast_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t n) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  const int64_t groups = n / 8;
  for (int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; g < groups;
       g += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float v[8];
    load8(src + g * 8, v);
    store8(dst + g * 8, v);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const int64_t i = groups * 8 + threadIdx.x;
    dst[i] = __float2bfloat16(src[i]);
  }
}


// ------------------------------------------------------------------------------------------
// row gather (a10, PPOAgent.update): dst[i, :] = src[idx[i], :] for rows of row_bytes (multiple of 16).  src may be
// PINNED HOST memory (unified addressing): the kernel then IS the host -> device transfer of the minibatch — only the rows
// of this minibatch cross PCIe, in permutation order, while the previous optimizer step still computes.  Measured
// (scripts/gpu_gather_bw.py, gpu_gather_overlap.py): 51 GB/s alone (cudaMemcpyAsync of the same bytes: 55 GB/s) with as
// few as 8 CTAs of 512 threads, four 16-byte loads per thread in flight.  Beside the PPO step it costs that step ~1 ms
// whatever the grid: the 384-thread GEMM variants use the whole register file (168 x 384), so an SM that hosts a gather
// CTA runs its GEMM CTA as a second wave — hence FEW fat CTAs for a SHORT time (callers pass max_ctas = 16) rather than
// one thin CTA per SM (148 x 128 threads: +4.4 ms on the step).
// ------------------------------------------------------------------------------------------
constexpr int kGatherThreads = 512;
__global__ void __launch_bounds__(kGatherThreads)
gather_rows_kernel(const uint4* __restrict__ src, const int64_t* __restrict__ idx, uint4* __restrict__ dst,
                   int64_t n_rows, int64_t units_per_row, int64_t src_rows) {
  pdl_sync();
  const int64_t total = n_rows * units_per_row;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t u0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; u0 < total; u0 += 4 * stride) {
    uint4 v[4];
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t u = u0 + j * stride;
      ok[j] = u < total;
      if (ok[j]) {
        const int64_t r = u / units_per_row, c = u - r * units_per_row;
        const int64_t sr = idx ? idx[r] : r;
        ok[j] = sr >= 0 && sr < src_rows;                          // an out-of-range index leaves the row untouched
        if (ok[j]) asm volatile("ld.global.cs.v4.u32 {%0, %1, %2, %3}, [%4];"
                                : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w)
                                : "l"(src + sr * units_per_row + c));
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (ok[j]) dst[u0 + j * stride] = v[j];
  }
}

}  // namespace

extern "C" {

int sc_patchify_bf16(const void* img, int img_is_bf16, void* out, int64_t B, int C, int H, int W, int P,
                     const int64_t* row_idx, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(img && out, SC_E_SHAPE, "sc_patchify_bf16: null pointer");
  SC_REQUIRE(B > 0 && C > 0 && P > 0 && P % 8 == 0 && H % P == 0 && W % P == 0, SC_E_SHAPE,
             "sc_patchify_bf16: need P %% 8 == 0 and H, W multiples of P (B=%lld C=%d H=%d W=%d P=%d)",
             (long long)B, C, H, W, P);
  SC_REQUIRE(sc_aligned16(img) && sc_aligned16(out), SC_E_ALIGN, "sc_patchify_bf16: unaligned pointer");
  const int64_t groups = B * C * H * (int64_t)W / 8;
  // the model's observation shapes: compile-time geometry (camera 3 x 128 x 128 / P 16, LiDAR-BEV 1 x 64 x 64 / P 8)
  {
    const int64_t token_rows = B * (H / P);
    const int grid_f = persistent_blocks(token_rows, 8);      // (BEV: 4 token rows per CTA iteration; the grid-stride loop covers the rest)
    cudaStream_t st = (cudaStream_t)stream;
    bool done = true;
    if (C == 3 && H == 128 && W == 128 && P == 16) {
      if (img_is_bf16) sc_launch_kernel(patchify_fixed_kernel<bf16, 3, 128, 128, 16>, dim3(grid_f), dim3(kThreads), 0, st, (const bf16*)img, (bf16*)out, token_rows, row_idx);
      else sc_launch_kernel(patchify_fixed_kernel<float, 3, 128, 128, 16>, dim3(grid_f), dim3(kThreads), 0, st, (const float*)img, (bf16*)out, token_rows, row_idx);
    } else if (C == 1 && H == 64 && W == 64 && P == 8) {
      if (img_is_bf16) sc_launch_kernel(patchify_fixed_kernel<bf16, 1, 64, 64, 8>, dim3(grid_f), dim3(kThreads), 0, st, (const bf16*)img, (bf16*)out, token_rows, row_idx);
      else sc_launch_kernel(patchify_fixed_kernel<float, 1, 64, 64, 8>, dim3(grid_f), dim3(kThreads), 0, st, (const float*)img, (bf16*)out, token_rows, row_idx);
    } else {
      done = false;
    }
    if (done) {
      sc_count_launch(1);
      SC_CHECK_LAUNCH("sc_patchify_bf16");
      return SC_OK;
    }
  }
  const int grid = persistent_blocks((groups + kThreads - 1) / kThreads, 8);
  if (img_is_bf16)
    sc_launch_kernel(patchify_kernel<bf16>, dim3(grid), dim3(kThreads), 0, (cudaStream_t)stream, (const bf16*)img, (bf16*)out, groups, C, H, W, P, row_idx);
  else
    sc_launch_kernel(patchify_kernel<float>, dim3(grid), dim3(kThreads), 0, (cudaStream_t)stream, (const float*)img, (bf16*)out, groups, C, H, W, P, row_idx);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_patchify_bf16");
  return SC_OK;
}

int sc_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                     int64_t M, int D, float eps, int io_f32, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(x && gamma && beta && y, SC_E_SHAPE, "sc_layernorm_fwd: null pointer");
  SC_REQUIRE(M > 0 && (D == 256 || D == 512 || D == 1024), SC_E_SHAPE,
             "sc_layernorm_fwd: M=%lld D=%d (D must be 256, 512 or 1024)", (long long)M, D);
  SC_REQUIRE(sc_aligned16(x) && sc_aligned16(y) && sc_aligned16(gamma) && sc_aligned16(beta), SC_E_ALIGN,
             "sc_layernorm_fwd: unaligned pointer");
  const int grid = persistent_blocks((M + 8 * 4 - 1) / (8 * 4), 3);
  cudaStream_t st = (cudaStream_t)stream;
#define LN_FWD(T, NCH, R) \
  sc_launch_kernel(layernorm_fwd_kernel<T, NCH, R>, dim3(grid), dim3(kThreads), 0, st, (const T*)x, gamma, beta, (T*)y, mean, rstd, M, eps)
  if (io_f32) {
    if (D == 256) LN_FWD(float, 1, 4); else if (D == 512) LN_FWD(float, 2, 1); else LN_FWD(float, 4, 1);
  } else {
    if (D == 256) LN_FWD(bf16, 1, 8); else if (D == 512) LN_FWD(bf16, 2, 2); else LN_FWD(bf16, 4, 1);
  }
#undef LN_FWD
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_layernorm_fwd");
  return SC_OK;
}

int sc_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                     const void* dres, void* dx, float* dgamma, float* dbeta, float* colsum, int64_t M, int D,
                     int io_f32, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(dy && x && mean && rstd && gamma && dx && dgamma && dbeta, SC_E_SHAPE, "sc_layernorm_bwd: null pointer");
  SC_REQUIRE(M > 0 && (D == 256 || D == 512 || D == 1024), SC_E_SHAPE,
             "sc_layernorm_bwd: M=%lld D=%d (D must be 256, 512 or 1024)", (long long)M, D);
  SC_REQUIRE(sc_aligned16(dy) && sc_aligned16(x) && sc_aligned16(dx) && sc_aligned16(dres) && sc_aligned16(gamma),
             SC_E_ALIGN, "sc_layernorm_bwd: unaligned pointer");
  const int grid = persistent_blocks((M + 8 * 4 - 1) / (8 * 4), 2);
  cudaStream_t st = (cudaStream_t)stream;
  const int reverse = sc_take_direction();
#define LN_BWD(T, NCH, R)                                                                                  \
  sc_launch_kernel(layernorm_bwd_kernel<T, NCH, R>, dim3(grid), dim3(kThreads), 0, st, (const T*)dy, (const T*)x, mean, rstd, gamma, \
                                                             (const T*)dres, (T*)dx, dgamma, dbeta, colsum, M, reverse)
  if (io_f32) {
    if (D == 256) LN_BWD(float, 1, 2); else if (D == 512) LN_BWD(float, 2, 1); else LN_BWD(float, 4, 1);
  } else {
    if (D == 256) LN_BWD(bf16, 1, 4); else if (D == 512) LN_BWD(bf16, 2, 1); else LN_BWD(bf16, 4, 1);
  }
#undef LN_BWD
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_layernorm_bwd");
  return SC_OK;
}

int sc_colsum_bf16(const void* x, float* out, int64_t M, int N, int64_t ld, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(x && out, SC_E_SHAPE, "sc_colsum_bf16: null pointer");
  SC_REQUIRE(M > 0 && N > 0 && N % 8 == 0 && ld >= N && ld % 8 == 0, SC_E_SHAPE,
             "sc_colsum_bf16: M=%lld N=%d ld=%lld (need N %% 8 == 0)", (long long)M, N, (long long)ld);
  SC_REQUIRE(sc_aligned16(x), SC_E_ALIGN, "sc_colsum_bf16: unaligned pointer");
  const int ychunks = (N + 1023) / 1024;
  const int target = (sc_num_sms() * 4 + ychunks - 1) / ychunks;
  int64_t rpb = (M + target - 1) / target;
  if (rpb < 32) rpb = 32;
  const dim3 grid((unsigned)((M + rpb - 1) / rpb), (unsigned)ychunks);
  sc_launch_kernel(colsum_kernel, dim3(grid), dim3(kThreads), 1024 * sizeof(float), (cudaStream_t)stream, (const bf16*)x, out, M, N, ld, rpb);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_colsum_bf16");
  return SC_OK;
}

int sc_meanpool_fwd(const void* x, void* out_bf16, float* out_f32, int64_t B, int S, int D, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(x && (out_bf16 || out_f32), SC_E_SHAPE, "sc_meanpool_fwd: null pointer");
  SC_REQUIRE(B > 0 && S > 0 && D % 256 == 0 && D <= 1024, SC_E_SHAPE, "sc_meanpool_fwd: B=%lld S=%d D=%d",
             (long long)B, S, D);
  SC_REQUIRE(sc_aligned16(x), SC_E_ALIGN, "sc_meanpool_fwd: unaligned pointer");
  const int grid = persistent_blocks(B, 16);
  sc_launch_kernel(meanpool_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, (const bf16*)x, (bf16*)out_bf16, out_f32, B, S, D);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_meanpool_fwd");
  return SC_OK;
}

int sc_mse_fwd_bwd(const void* r, const void* t, const void* t1, int seg_rows, int D, float* loss_accum, void* dr,
                   int64_t numel, float inv_numel, float grad_scale, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(r && t && loss_accum, SC_E_SHAPE, "sc_mse_fwd_bwd: null pointer");
  SC_REQUIRE(numel > 0 && numel % 8 == 0, SC_E_SHAPE, "sc_mse_fwd_bwd: numel=%lld must be a positive multiple of 8",
             (long long)numel);
  SC_REQUIRE(sc_aligned16(r) && sc_aligned16(t) && sc_aligned16(t1) && sc_aligned16(dr), SC_E_ALIGN,
             "sc_mse_fwd_bwd: unaligned pointer");
  SC_REQUIRE(!t1 || (seg_rows > 0 && D > 0 && D % 8 == 0 && numel % (2LL * seg_rows * D) == 0), SC_E_SHAPE,
             "sc_mse_fwd_bwd: two-target mode needs numel %% (2*seg_rows*D) == 0");
  const int64_t groups = numel / 8;
  const int grid = persistent_blocks((groups + kThreads - 1) / kThreads, 8);
  sc_launch_kernel(mse_kernel, dim3(grid), dim3(kThreads), 0, (cudaStream_t)stream, (const bf16*)r, (const bf16*)t, (const bf16*)t1, loss_accum,
                                                          (bf16*)dr, groups, inv_numel, grad_scale, seg_rows,
                                                          D > 0 ? D / 8 : 1);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_mse_fwd_bwd");
  return SC_OK;
}

int sc_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(src && dst && n > 0, SC_E_SHAPE, "sc_cast_f32_bf16: bad arguments");
  SC_REQUIRE(sc_aligned16(src) && sc_aligned16(dst), SC_E_ALIGN, "sc_cast_f32_bf16: unaligned pointer");
  const int grid = persistent_blocks((n / 8 + kThreads - 1) / kThreads, 8);
  sc_launch_kernel(cast_kernel, dim3(grid), dim3(kThreads), 0, (cudaStream_t)stream, src, (bf16*)dst, n);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_cast_f32_bf16");
  return SC_OK;
}

int sc_gather_rows(const void* src, int64_t src_rows, int64_t row_bytes, const int64_t* idx, void* dst, int64_t n_rows,
                   int max_ctas, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(src && dst && n_rows > 0 && src_rows > 0 && row_bytes > 0 && row_bytes % 16 == 0, SC_E_SHAPE,
             "sc_gather_rows: n_rows=%lld src_rows=%lld row_bytes=%lld", (long long)n_rows, (long long)src_rows,
             (long long)row_bytes);
  SC_REQUIRE(sc_aligned16(src) && sc_aligned16(dst), SC_E_ALIGN, "sc_gather_rows: unaligned pointer");
  const int64_t units = row_bytes / 16;
  int64_t blocks = (n_rows * units + 4 * kGatherThreads - 1) / (4 * kGatherThreads);
  const int64_t cap = max_ctas > 0 ? max_ctas : sc_num_sms();
  if (blocks > cap) blocks = cap;
  sc_launch_kernel(gather_rows_kernel, dim3((unsigned)blocks), dim3(kGatherThreads), 0, (cudaStream_t)stream,
                   static_cast<const uint4*>(src), idx, static_cast<uint4*>(dst), n_rows, units, src_rows);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_gather_rows");
  return SC_OK;
}

}  // extern "C"
