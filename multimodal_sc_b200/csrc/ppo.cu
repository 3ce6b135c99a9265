// ppo.cu — actor-critic output layer, action sampling, GAE reverse scan, PPO clipped loss
// (forward + backward fused) and the fused flat Adam.  SPEC v0 rows a7, a8, a9, a10.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxA1 = 16;     // actions + 1 (value) supported by the register-resident heads

__device__ __forceinline__ void load8f(const bf16* p, float (&x)[8]) {
  uint4 w = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16(w.x), b = unpack_bf16(w.y), c = unpack_bf16(w.z), d = unpack_bf16(w.w);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
}
__device__ __forceinline__ void load8f(const float* p, float (&x)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}

// ------------------------------------------------------------------------------------------
// heads: warp per sample, lane owns 8 trunk features (D = 256); A1 <= 16 outputs
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
heads_fwd_kernel(const bf16* __restrict__ trunk, const float* __restrict__ W, const float* __restrict__ bias,
                 float* __restrict__ out, int64_t B, int D, int A1) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t b = warp0; b < B; b += nwarps) {
    float acc[kMaxA1];
#pragma unroll
    for (int j = 0; j < kMaxA1; ++j) acc[j] = 0.f;
    for (int c = 0; c < D; c += 256) {
      float t[8];
      load8f(trunk + b * D + c + lane * 8, t);
#pragma unroll
      for (int j = 0; j < kMaxA1; ++j) {
        if (j < A1) {
          float w[8];
          load8f(W + static_cast<int64_t>(j) * D + c + lane * 8, w);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[j] = fmaf(t[i], w[i], acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kMaxA1; ++j) {
      if (j < A1) {
        const float s = warp_sum(acc[j]);
        if (lane == 0) out[b * A1 + j] = s + bias[j];
      }
    }
  }
}

// dtrunk_pre = (dho W) (1 - trunk^2); dW += dho^T trunk; dbias += colsum(dho).  D == 256.
__global__ void __launch_bounds__(kThreads)
heads_bwd_kernel(const float* __restrict__ dho, const bf16* __restrict__ trunk, const float* __restrict__ W,
                 bf16* __restrict__ dtrunk_pre, float* __restrict__ dW, float* __restrict__ dbias, int64_t B,
                 int A1) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  constexpr int D = 256;
  __shared__ float sW[kMaxA1 * D];
  __shared__ float sdW[kMaxA1 * D];
  __shared__ float sdb[kMaxA1];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < A1 * D; i += blockDim.x) { sW[i] = W[i]; sdW[i] = 0.f; }
  if (threadIdx.x < kMaxA1) sdb[threadIdx.x] = 0.f;
  __syncthreads();
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  float adw[kMaxA1][8];
  float adb = 0.f;                                   // lane j accumulates dbias[j]
#pragma unroll
  for (int j = 0; j < kMaxA1; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) adw[j][i] = 0.f;
  for (int64_t b = warp0; b < B; b += nwarps) {
    float t[8], dt[8];
    load8f(trunk + b * D + lane * 8, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) dt[i] = 0.f;
    const float mine = lane < A1 ? dho[b * A1 + lane] : 0.f;
    adb += mine;
#pragma unroll
    for (int j = 0; j < kMaxA1; ++j) {
      if (j < A1) {
        const float g = __shfl_sync(0xffffffffu, mine, j);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          dt[i] = fmaf(g, sW[j * D + lane * 8 + i], dt[i]);
          adw[j][i] = fmaf(g, t[i], adw[j][i]);
        }
      }
    }
    uint4 o = make_uint4(pack_bf16(dt[0] * (1.f - t[0] * t[0]), dt[1] * (1.f - t[1] * t[1])),
                         pack_bf16(dt[2] * (1.f - t[2] * t[2]), dt[3] * (1.f - t[3] * t[3])),
                         pack_bf16(dt[4] * (1.f - t[4] * t[4]), dt[5] * (1.f - t[5] * t[5])),
                         pack_bf16(dt[6] * (1.f - t[6] * t[6]), dt[7] * (1.f - t[7] * t[7])));
    *reinterpret_cast<uint4*>(dtrunk_pre + b * D + lane * 8) = o;
  }
#pragma unroll
  for (int j = 0; j < kMaxA1; ++j) {
    if (j < A1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&sdW[j * D + lane * 8 + i], adw[j][i]);
    }
  }
  if (lane < A1) atomicAdd(&sdb[lane], adb);
  __syncthreads();
  for (int i = threadIdx.x; i < A1 * D; i += blockDim.x) atomicAdd(dW + i, sdW[i]);
  if (threadIdx.x < A1) atomicAdd(dbias + threadIdx.x, sdb[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------
// inverse-CDF sampling, fp64 cdf in index order (SPEC v0 a7); logp in fp32
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
sample_kernel(const float* __restrict__ ho, int ld, const float* __restrict__ u, int64_t* __restrict__ action,
              float* __restrict__ logp, int64_t B, int A) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  for (int64_t b = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; b < B;
       b += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float* l = ho + b * ld;
    float mx = l[0];
    for (int j = 1; j < A; ++j) mx = fmaxf(mx, l[j]);
    double den = 0.0;
    for (int j = 0; j < A; ++j) den += exp(static_cast<double>(l[j]) - static_cast<double>(mx));
    const double uu = static_cast<double>(u[b]);
    double cdf = 0.0;
    int a = 0;
    for (int j = 0; j < A; ++j) {
      cdf += exp(static_cast<double>(l[j]) - static_cast<double>(mx)) / den;
      a += (cdf <= uu) ? 1 : 0;
    }
    a = min(a, A - 1);
    float se = 0.f;
    for (int j = 0; j < A; ++j) se += expf(l[j] - mx);
    action[b] = a;
    logp[b] = (l[a] - mx) - logf(se);
  }
}

// ------------------------------------------------------------------------------------------
// GAE: A_t = delta_t + c_t A_{t+1} is a composition of affine maps f_t(x) = delta_t + c_t x.
// Block = 32 warps = 32 environments; a [TC x 32] tile of r, v, d is loaded coalesced into
// smem, then warp w scans environment w: lane l holds timestep l of each 32-step segment and
// a suffix scan of (a, b) pairs with __shfl_down composes the maps; segments are chained from
// the end of the rollout backwards with the carried A_{t+1}.
// ------------------------------------------------------------------------------------------
constexpr int kGaeTC = 64;
__global__ void __launch_bounds__(1024)
gae_kernel(const float* __restrict__ rew, const float* __restrict__ val, const float* __restrict__ done,
           const float* __restrict__ last_value, float* __restrict__ adv, float* __restrict__ ret,
           double* __restrict__ sums, int T, int64_t N, float gamma, float lam) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  __shared__ float sr[kGaeTC][33], sv[kGaeTC + 1][33], sd[kGaeTC][33], sa[kGaeTC][33];
  __shared__ double ssum[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * 32;
  if (threadIdx.x < 2) ssum[threadIdx.x] = 0.0;
  const int64_t n_col = n0 + lane;                   // column this lane loads/stores
  const int64_t n_env = n0 + warp;                   // environment this warp scans
  float carry_a = 0.f;                               // A_{t+1} entering the current chunk
  double acc_s = 0.0, acc_q = 0.0;
  for (int t_hi = T; t_hi > 0; t_hi -= kGaeTC) {
    const int t_lo = max(0, t_hi - kGaeTC);
    const int tc = t_hi - t_lo;
    __syncthreads();
    for (int t = warp; t < tc; t += 32) {
      const bool ok = n_col < N;
      const int64_t o = static_cast<int64_t>(t_lo + t) * N + n_col;
      sr[t][lane] = ok ? rew[o] : 0.f;
      sv[t][lane] = ok ? val[o] : 0.f;
      sd[t][lane] = ok ? done[o] : 0.f;
    }
    if (warp == 0) {                                 // V_{t_hi}: next chunk's first value or last_value
      float vn = 0.f;
      if (n_col < N) vn = (t_hi == T) ? last_value[n_col] : val[static_cast<int64_t>(t_hi) * N + n_col];
      sv[tc][lane] = vn;
    }
    __syncthreads();
    // scan segments of 32 timesteps from the end of the chunk backwards
    for (int seg_hi = tc; seg_hi > 0; seg_hi -= 32) {
      const int seg_lo = max(0, seg_hi - 32);
      const int t = seg_lo + lane;
      float a = 1.f, b = 0.f;                        // identity map for lanes past the segment
      if (t < seg_hi) {
        const float nd = 1.f - sd[t][warp];
        b = sr[t][warp] + gamma * sv[t + 1][warp] * nd - sv[t][warp];      // delta_t
        a = gamma * lam * nd;                                               // c_t
      }
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const float a2 = __shfl_down_sync(0xffffffffu, a, off);
        const float b2 = __shfl_down_sync(0xffffffffu, b, off);
        if (lane + off < 32) { b = fmaf(a, b2, b); a *= a2; }
      }
      const float A = fmaf(a, carry_a, b);
      if (t < seg_hi) sa[t][warp] = A;
      carry_a = __shfl_sync(0xffffffffu, A, 0);
    }
    __syncthreads();
    for (int t = warp; t < tc; t += 32) {
      if (n_col < N) {
        const int64_t o = static_cast<int64_t>(t_lo + t) * N + n_col;
        const float A = sa[t][lane];
        adv[o] = A;
        ret[o] = A + sv[t][lane];
        acc_s += A;
        acc_q += static_cast<double>(A) * A;
      }
    }
  }
  (void)n_env;
  if (sums) {
    for (int o = 16; o > 0; o >>= 1) {
      acc_s += __shfl_xor_sync(0xffffffffu, acc_s, o);
      acc_q += __shfl_xor_sync(0xffffffffu, acc_q, o);
    }
    if (lane == 0) { atomicAdd(&ssum[0], acc_s); atomicAdd(&ssum[1], acc_q); }
    __syncthreads();
    if (threadIdx.x < 2) atomicAdd(sums + threadIdx.x, ssum[threadIdx.x]);
  }
}

// Large-N form: one thread per environment walks t = T-1 .. 0 (the recurrence is a single FMA per step);
// loads of 8 timesteps are issued before they are consumed, every access is coalesced across environments.
// Same arithmetic as the oracle's loop.  Used when N is large enough to fill the machine this way.
__global__ void __launch_bounds__(128)
gae_thread_kernel(const float* __restrict__ rew, const float* __restrict__ val, const float* __restrict__ done,
                  const float* __restrict__ last_value, float* __restrict__ adv, float* __restrict__ ret,
                  double* __restrict__ sums, int T, int64_t N, float gamma, float lam) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  __shared__ double ssum[2];
  if (threadIdx.x < 2) ssum[threadIdx.x] = 0.0;
  __syncthreads();
  const int64_t n = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  double acc_s = 0.0, acc_q = 0.0;
  if (n < N) {
    float next_v = last_value[n], A = 0.f;
    constexpr int U = 8;
    for (int t_hi = T; t_hi > 0; t_hi -= U) {
      float r[U], v[U], d[U];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int t = t_hi - 1 - j;
        if (t >= 0) {
          const int64_t o = static_cast<int64_t>(t) * N + n;
          r[j] = rew[o]; v[j] = val[o]; d[j] = done[o];
        }
      }
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int t = t_hi - 1 - j;
        if (t >= 0) {
          const float nd = 1.f - d[j];
          const float delta = r[j] + gamma * next_v * nd - v[j];
          A = delta + gamma * lam * nd * A;
          const int64_t o = static_cast<int64_t>(t) * N + n;
          adv[o] = A;
          ret[o] = A + v[j];
          acc_s += A;
          acc_q += static_cast<double>(A) * A;
          next_v = v[j];
        }
      }
    }
  }
  if (sums) {
    for (int o = 16; o > 0; o >>= 1) {
      acc_s += __shfl_xor_sync(0xffffffffu, acc_s, o);
      acc_q += __shfl_xor_sync(0xffffffffu, acc_q, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&ssum[0], acc_s); atomicAdd(&ssum[1], acc_q); }
    __syncthreads();
    if (threadIdx.x < 2) atomicAdd(sums + threadIdx.x, ssum[threadIdx.x]);
  }
}

__global__ void __launch_bounds__(kThreads)
adv_normalize_kernel(float* __restrict__ adv, const double* __restrict__ sums, int64_t count, int64_t n,
                     float eps) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  const double mean = sums[0] / static_cast<double>(count);
  double var = sums[1] / static_cast<double>(count) - mean * mean;
  var = var > 0.0 ? var : 0.0;
  const float m = static_cast<float>(mean);
  const float inv = 1.0f / (static_cast<float>(sqrt(var)) + eps);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    adv[i] = (adv[i] - m) * inv;
}

// ------------------------------------------------------------------------------------------
// PPO clipped loss, forward + backward fused: one thread per sample, block reduce of 5 stats
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
ppo_loss_kernel(const float* __restrict__ ho, int ld, const int64_t* __restrict__ actions,
                const float* __restrict__ logp_old, const float* __restrict__ adv, const float* __restrict__ ret,
                const int64_t* __restrict__ idx, float* __restrict__ dho, float* __restrict__ stats, int64_t M, int A,
                float clip_eps, float c_v, float c_ent, float inv_m) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  __shared__ float red[5][kThreads / 32];
  float st[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < M;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float* l = ho + i * ld;
    float lg[kMaxA1];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kMaxA1; ++j)
      if (j < A) { lg[j] = l[j]; mx = fmaxf(mx, lg[j]); }
    float se = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxA1; ++j)
      if (j < A) se += expf(lg[j] - mx);
    const float lse = mx + logf(se);
    const int64_t src = idx ? idx[i] : i;                      // fused minibatch gather of the rollout vectors
    const int a = static_cast<int>(actions[src]);
    // an action outside [0, A) poisons the statistics instead of silently contributing logp = 0
    float ent = 0.f, logp = (a >= 0 && a < A) ? 0.f : __int_as_float(0x7fc00000);
    float p[kMaxA1], lp[kMaxA1];
#pragma unroll
    for (int j = 0; j < kMaxA1; ++j)
      if (j < A) {
        lp[j] = lg[j] - lse;
        p[j] = expf(lp[j]);
        ent -= p[j] * lp[j];
        if (j == a) logp = lp[j];
      }
    const float Ad = adv[src];
    const float lp_old = logp_old[src];
    const float ratio = expf(logp - lp_old);
    const float s1 = ratio * Ad;
    const float rc = fminf(fmaxf(ratio, 1.f - clip_eps), 1.f + clip_eps);
    const float s2 = rc * Ad;
    const bool inside = (ratio >= 1.f - clip_eps) && (ratio <= 1.f + clip_eps);
    const float gmin = (inside || s1 < s2) ? Ad : 0.f;          // d min(s1,s2) / d ratio
    const float v = l[A];
    const float dv = v - ret[src];
    st[0] += -fminf(s1, s2);
    st[1] += 0.5f * dv * dv;
    st[2] += ent;
    st[3] += lp_old - logp;
    st[4] += (fabsf(ratio - 1.f) > clip_eps) ? 1.f : 0.f;
    // d total / d logit_j = inv_m * [ -gmin*ratio*(1[j==a] - p_j) + c_ent * p_j (lp_j + ent) ]
    const float gp = -gmin * ratio;
    float* d = dho + i * ld;
#pragma unroll
    for (int j = 0; j < kMaxA1; ++j)
      if (j < A) d[j] = inv_m * (gp * ((j == a ? 1.f : 0.f) - p[j]) + c_ent * p[j] * (lp[j] + ent));
    d[A] = inv_m * c_v * dv;
  }
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const float w = warp_sum(st[s]);
    if ((threadIdx.x & 31) == 0) red[s][threadIdx.x >> 5] = w;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += red[threadIdx.x][w];
    atomicAdd(stats + threadIdx.x, t * inv_m);
  }
}

// ------------------------------------------------------------------------------------------
// step_begin: zero the flat gradient bucket and the statistics, advance the device-side step counter
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
step_begin_kernel(float4* __restrict__ grad, int64_t n4, float* __restrict__ stats, int n_stats,
                  int64_t* __restrict__ step_dev) {
  pdl_sync();
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  for (int64_t i = t; i < n4; i += static_cast<int64_t>(gridDim.x) * blockDim.x) grad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t < n_stats) stats[t] = 0.f;
  if (t == 0 && step_dev) *step_dev += 1;
}

// ------------------------------------------------------------------------------------------
// SNR embedding (SNR-conditioned encoders): h = tanh(s u + v), emb = W h
// fwd: one CTA per sample (h in shared memory, thread d owns emb[b, d]).
// bwd: CTA per 8-sample group computes dh[b, :] = (demb[b, :] W) (1 - h^2) in shared memory and accumulates
//      du / dv; dW[d, e] += sum_b demb[b, d] h[b, e] by a second kernel with one thread per (d, e).
// ------------------------------------------------------------------------------------------
constexpr int kSnrMaxE = 256;

__global__ void __launch_bounds__(kThreads)
snr_embed_fwd_kernel(const float* __restrict__ snr_db, float snr_scalar, const float* __restrict__ u,
                     const float* __restrict__ v, const float* __restrict__ W, float* __restrict__ emb,
                     float* __restrict__ h, int64_t B, int D, int E) {
  pdl_sync();
  __shared__ float hs[kSnrMaxE];
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const float s = (snr_db ? snr_db[b] : snr_scalar) * 0.05f;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      const float t = tanhf(fmaf(s, u[e], v[e]));
      hs[e] = t;
      h[b * E + e] = t;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      const float* w = W + static_cast<int64_t>(d) * E;
      float acc = 0.f;
      for (int e = 0; e < E; ++e) acc = fmaf(w[e], hs[e], acc);
      emb[b * D + d] = acc;
    }
    __syncthreads();
  }
}

// du[e] += sum_b dh[b, e] s_b, dv[e] += sum_b dh[b, e]; dh = (demb W) (1 - h^2).  One CTA per sample group.
__global__ void __launch_bounds__(kThreads)
snr_embed_bwd_uv_kernel(const float* __restrict__ demb, const float* __restrict__ h, const float* __restrict__ snr_db,
                        float snr_scalar, const float* __restrict__ W, float* __restrict__ du, float* __restrict__ dv,
                        float demb_scale, int64_t B, int D, int E) {
  pdl_sync();
  __shared__ float ds[1024];
  const int e = threadIdx.x;
  float au = 0.f, av = 0.f;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    for (int d = threadIdx.x; d < D; d += blockDim.x) ds[d] = demb[b * D + d] * demb_scale;
    __syncthreads();
    if (e < E) {
      float acc = 0.f;
      for (int d = 0; d < D; ++d) acc = fmaf(ds[d], W[static_cast<int64_t>(d) * E + e], acc);
      const float t = h[b * E + e];
      const float dh = acc * (1.0f - t * t);
      const float s = (snr_db ? snr_db[b] : snr_scalar) * 0.05f;
      au = fmaf(dh, s, au);
      av += dh;
    }
    __syncthreads();
  }
  if (e < E) {
    atomicAdd(du + e, au);
    atomicAdd(dv + e, av);
  }
}

// dW[d, e] += sum_b demb[b, d] h[b, e]: one thread per (d, e), the batch split over gridDim.y
__global__ void __launch_bounds__(kThreads)
snr_embed_bwd_w_kernel(const float* __restrict__ demb, const float* __restrict__ h, float* __restrict__ dW,
                       float demb_scale, int64_t B, int D, int E) {
  pdl_sync();
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (t >= static_cast<int64_t>(D) * E) return;
  const int d = static_cast<int>(t / E), e = static_cast<int>(t % E);
  const int64_t per = (B + gridDim.y - 1) / gridDim.y;
  const int64_t b0 = blockIdx.y * per, b1 = min(B, b0 + per);
  float acc = 0.f;
  for (int64_t b = b0; b < b1; ++b) acc = fmaf(demb[b * D + d], h[b * E + e], acc);
  atomicAdd(dW + t, acc * demb_scale);
}

// ------------------------------------------------------------------------------------------
// fused flat Adam (torch.optim.Adam arithmetic order) + bf16 shadow refresh
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            bf16* __restrict__ pb, int64_t n, float step_size, float b1, float b2, float eps, float inv_sqrt_bc2,
            float grad_scale, float lr, int64_t step, const int64_t* __restrict__ step_dev) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  if (step_dev) {
    // graph replay: the step counter lives on the device; one thread per CTA redoes the host's double-precision
    // bias corrections
    __shared__ float bc[2];
    if (threadIdx.x == 0) {
      adam_bias_corrections(static_cast<double>(step + *step_dev), lr, b1, b2, bc[0], bc[1]);
    }
    __syncthreads();
    step_size = bc[0];
    inv_sqrt_bc2 = bc[1];
  }
  const int64_t n4 = n / 4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* P = &pp.x; float* G = &gg.x; float* Mo = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      adam_update(P[j], G[j] * grad_scale, Mo[j], V[j], step_size, b1, b2, eps, inv_sqrt_bc2);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (pb) reinterpret_cast<uint2*>(pb)[i] = make_uint2(pack_bf16(pp.x, pp.y), pack_bf16(pp.z, pp.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = n4 * 4 + threadIdx.x;
    float mo = m[i], vo = v[i], po = p[i];
    adam_update(po, g[i] * grad_scale, mo, vo, step_size, b1, b2, eps, inv_sqrt_bc2);
    m[i] = mo; v[i] = vo; p[i] = po;
    if (pb) pb[i] = __float2bfloat16(po);
  }
}

inline int grid_for(int64_t work_items, int per_block, int per_sm) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  int64_t cap = static_cast<int64_t>(sc_num_sms()) * per_sm;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks < cap ? blocks : cap);
}

}  // namespace

extern "C" {

int sc_heads_fwd(const void* trunk, const float* W, const float* bias, float* head_out, int64_t B, int D, int A1,
                 void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(trunk && W && bias && head_out, SC_E_SHAPE, "sc_heads_fwd: null pointer");
  SC_REQUIRE(B > 0 && D % 256 == 0 && A1 >= 2 && A1 <= kMaxA1, SC_E_SHAPE, "sc_heads_fwd: B=%lld D=%d A1=%d",
             (long long)B, D, A1);
  SC_REQUIRE(sc_aligned16(trunk) && sc_aligned16(W), SC_E_ALIGN, "sc_heads_fwd: unaligned pointer");
  sc_launch_kernel(heads_fwd_kernel, dim3(grid_for(B, 8, 8)), dim3(kThreads), 0, (cudaStream_t)stream, (const bf16*)trunk, W, bias, head_out,
                                                                            B, D, A1);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_heads_fwd");
  return SC_OK;
}

int sc_heads_bwd(const float* dho, const void* trunk, const float* W, void* dtrunk_pre, float* dW, float* dbias,
                 int64_t B, int D, int A1, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(dho && trunk && W && dtrunk_pre && dW && dbias, SC_E_SHAPE, "sc_heads_bwd: null pointer");
  SC_REQUIRE(B > 0 && D == 256 && A1 >= 2 && A1 <= kMaxA1, SC_E_SHAPE, "sc_heads_bwd: B=%lld D=%d A1=%d (D must be 256)",
             (long long)B, D, A1);
  SC_REQUIRE(sc_aligned16(trunk) && sc_aligned16(dtrunk_pre), SC_E_ALIGN, "sc_heads_bwd: unaligned pointer");
  sc_launch_kernel(heads_bwd_kernel, dim3(grid_for(B, 64, 1)), dim3(kThreads), 0, (cudaStream_t)stream, dho, (const bf16*)trunk, W,
                                                                             (bf16*)dtrunk_pre, dW, dbias, B, A1);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_heads_bwd");
  return SC_OK;
}

int sc_sample_action(const float* head_out, int ld, const float* u, int64_t* action, float* logp, int64_t B, int A,
                     void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(head_out && u && action && logp, SC_E_SHAPE, "sc_sample_action: null pointer");
  SC_REQUIRE(B > 0 && A >= 1 && ld >= A, SC_E_SHAPE, "sc_sample_action: B=%lld A=%d ld=%d", (long long)B, A, ld);
  sc_launch_kernel(sample_kernel, dim3(grid_for(B, kThreads, 8)), dim3(kThreads), 0, (cudaStream_t)stream, head_out, ld, u, action, logp, B, A);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_sample_action");
  return SC_OK;
}

int sc_gae(const float* rewards, const float* values, const float* dones, const float* last_value, float* adv,
           float* ret, double* sums, int T, int64_t N, float gamma, float lam, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(rewards && values && dones && last_value && adv && ret, SC_E_SHAPE, "sc_gae: null pointer");
  SC_REQUIRE(T > 0 && N > 0, SC_E_SHAPE, "sc_gae: T=%d N=%lld", T, (long long)N);
  if (N >= 16384) {
    // enough environments to fill the GPU with one thread each: fully coalesced streaming form
    sc_launch_kernel(gae_thread_kernel, dim3((int)((N + 127) / 128)), dim3(128), 0, (cudaStream_t)stream, rewards, values, dones, last_value,
                                                                               adv, ret, sums, T, N, gamma, lam);
  } else {
    // few environments (the PPO configs: 32..256): parallelise over time with the warp-shuffle scan
    const int grid = (int)((N + 31) / 32);
    sc_launch_kernel(gae_kernel, dim3(grid), dim3(1024), 0, (cudaStream_t)stream, rewards, values, dones, last_value, adv, ret, sums, T, N,
                                                        gamma, lam);
  }
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_gae");
  return SC_OK;
}

int sc_adv_normalize(float* adv, const double* sums, int64_t count, int64_t n, float eps, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(adv && sums && count > 0 && n > 0, SC_E_SHAPE, "sc_adv_normalize: bad arguments");
  sc_launch_kernel(adv_normalize_kernel, dim3(grid_for(n, kThreads, 8)), dim3(kThreads), 0, (cudaStream_t)stream, adv, sums, count, n, eps);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_adv_normalize");
  return SC_OK;
}

int sc_ppo_loss(const float* head_out, int ld, const int64_t* actions, const float* logp_old, const float* adv,
                const float* ret, const int64_t* idx, float* dho, float* stats, int64_t M, int A, float clip_eps,
                float c_v, float c_ent, float inv_m, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(head_out && actions && logp_old && adv && ret && dho && stats, SC_E_SHAPE, "sc_ppo_loss: null pointer");
  SC_REQUIRE(M > 0 && A >= 1 && A < kMaxA1 && ld >= A + 1, SC_E_SHAPE, "sc_ppo_loss: M=%lld A=%d ld=%d",
             (long long)M, A, ld);
  sc_launch_kernel(ppo_loss_kernel, dim3(grid_for(M, kThreads, 4)), dim3(kThreads), 0, (cudaStream_t)stream, 
      head_out, ld, actions, logp_old, adv, ret, idx, dho, stats, M, A, clip_eps, c_v, c_ent, inv_m);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_ppo_loss");
  return SC_OK;
}

int sc_step_begin(float* grad, int64_t n, float* stats, int n_stats, int64_t* step_dev, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(n >= 0 && n % 4 == 0 && n_stats >= 0 && n_stats <= kThreads && (n == 0 || grad) && (n_stats == 0 || stats),
             SC_E_SHAPE, "sc_step_begin: n=%lld (multiple of 4) n_stats=%d (<= %d)", (long long)n, n_stats, kThreads);
  SC_REQUIRE(sc_aligned16(grad), SC_E_ALIGN, "sc_step_begin: unaligned gradient bucket");
  sc_launch_kernel(step_begin_kernel, dim3(grid_for(n / 4 + 1, kThreads, 8)), dim3(kThreads), 0, (cudaStream_t)stream,
                   reinterpret_cast<float4*>(grad), n / 4, stats, n_stats, step_dev);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_step_begin");
  return SC_OK;
}

int sc_snr_embed_fwd(const float* snr_db, float snr_db_scalar, const float* u, const float* v, const float* W,
                     float* emb, float* h, int64_t B, int D, int E, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(u && v && W && emb && h, SC_E_SHAPE, "sc_snr_embed_fwd: null pointer");
  SC_REQUIRE(B > 0 && D > 0 && D <= 1024 && E > 0 && E <= kSnrMaxE, SC_E_SHAPE, "sc_snr_embed_fwd: B=%lld D=%d E=%d",
             (long long)B, D, E);
  const int grid = (int)(B < (int64_t)sc_num_sms() * 8 ? B : (int64_t)sc_num_sms() * 8);
  sc_launch_kernel(snr_embed_fwd_kernel, dim3(grid), dim3(kThreads), 0, (cudaStream_t)stream, snr_db, snr_db_scalar, u, v,
                   W, emb, h, B, D, E);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_snr_embed_fwd");
  return SC_OK;
}

int sc_snr_embed_bwd(const float* demb, const float* h, const float* snr_db, float snr_db_scalar, const float* W,
                     float* dW, float* du, float* dv, float demb_scale, int64_t B, int D, int E, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(demb && h && W && dW && du && dv, SC_E_SHAPE, "sc_snr_embed_bwd: null pointer");
  SC_REQUIRE(B > 0 && D > 0 && D <= 1024 && E > 0 && E <= kSnrMaxE && E <= kThreads, SC_E_SHAPE,
             "sc_snr_embed_bwd: B=%lld D=%d E=%d", (long long)B, D, E);
  const int g1 = (int)(B < (int64_t)sc_num_sms() * 2 ? B : (int64_t)sc_num_sms() * 2);
  sc_launch_kernel(snr_embed_bwd_uv_kernel, dim3(g1), dim3(kThreads), 0, (cudaStream_t)stream, demb, h, snr_db,
                   snr_db_scalar, W, du, dv, demb_scale, B, D, E);
  const int gx = (int)(((int64_t)D * E + kThreads - 1) / kThreads);
  int gy = (int)((B + 255) / 256);
  if (gy > 16) gy = 16;
  sc_launch_kernel(snr_embed_bwd_w_kernel, dim3(gx, gy), dim3(kThreads), 0, (cudaStream_t)stream, demb, h, dW, demb_scale, B, D, E);
  sc_count_launch(2);
  SC_CHECK_LAUNCH("sc_snr_embed_bwd");
  return SC_OK;
}

int sc_adam_flat(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1,
                 float beta2, float eps, int64_t step, float grad_scale, const int64_t* step_dev, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(p && g && m && v && n > 0 && (step >= 1 || (step_dev && step >= 0)), SC_E_SHAPE, "sc_adam_flat: bad arguments");
  SC_REQUIRE(sc_aligned16(p) && sc_aligned16(g) && sc_aligned16(m) && sc_aligned16(v) &&
                 (reinterpret_cast<uintptr_t>(p_bf16) & 7) == 0,
             SC_E_ALIGN, "sc_adam_flat: unaligned pointer");
  const double tstep = (double)(step >= 1 ? step : 1);
  const double bc1 = 1.0 - pow((double)beta1, tstep);
  const double bc2 = 1.0 - pow((double)beta2, tstep);
  const float step_size = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  sc_launch_kernel(adam_kernel, dim3(grid_for(n / 4 + 1, kThreads, 8)), dim3(kThreads), 0, (cudaStream_t)stream, 
      p, g, m, v, (bf16*)p_bf16, n, step_size, beta1, beta2, eps, inv_sqrt_bc2, grad_scale, lr, step, step_dev);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_adam_flat");
  return SC_OK;
}

}  // extern "C"
