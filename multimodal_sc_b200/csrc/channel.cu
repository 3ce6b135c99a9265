// channel.cu — fused power normalisation + noisy channel (SPEC v0 rows a3 + a4).
//
// One CTA per (sample, modality) row of k real symbols: one pass computes sum z^2 (warp shuffle
// + smem), the scale rsqrt(mean z^2 + eps), draws the noise in-kernel (Philox4x32-10 +
// Box-Muller, counter layout of oracle/philox.py) and writes y.  z stays in registers between
// the reduction and the output, so HBM traffic is the algorithmic 1 read + 1 write.
// Backward needs no noise at all (it is additive): dz = s (g - zn mean(g zn)).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxV = 4;                      // float4 per thread -> k <= 4096

#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
  // one IMAD.WIDE per multiplier per round (hi and lo halves of the 64-bit product) instead of IMAD.HI + IMAD;
  // k0 + r W0 / k1 + r W1 are loop-invariant in the callers' row loops (constant offsets of two registers)
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    if (r > 0) { k0 += PHILOX_W0; k1 += PHILOX_W1; }
    const uint64_t p0 = static_cast<uint64_t>(PHILOX_M0) * c.x;
    const uint64_t p1 = static_cast<uint64_t>(PHILOX_M1) * c.z;
    const uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
    const uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
  }
  return c;
}
__device__ __forceinline__ float u32_to_uniform(uint32_t x) {
  return (static_cast<float>(x >> 9) + 0.5f) * 1.1920928955078125e-07f;   // 2^-23, exact
}
// Reference Box-Muller with the accurate logf / sincospif (~60 instructions per pair): kept for the block-fading
// coefficient (one call per row) and as the A/B yardstick of the fast form (sc_debug_set("chan_exact", 1)).
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float r = sqrtf(-2.0f * logf(u32_to_uniform(a)));
  float s, c;
  sincospif(2.0f * u32_to_uniform(b), &s, &c);
  n0 = r * c;
  n1 = r * s;
}
// Fast Box-Muller (~19 instructions per pair) with a BOUNDED absolute error of the normals: worst case 6e-6 at the
// extreme tail (r = 5.8), ~1e-6 typical; measured 1.8e-6 over 2^26 draws (the accurate variant: 6.5e-7),
// tests/test_gpu_kernels.py::test_box_muller_fast_error_bound holds it to 8e-6.  The parity bar is fp32 rtol 1e-5 + atol 2e-5 on y = zn + sigma n with sigma <= 1.78.
//   u  = ((x >> 9) + 0.5) 2^-23 is built without an int->float conversion: f = 1.mantissa in [1, 2), u = f - (1 - 2^-24)
//        (exact: the result (2k + 1) 2^-24 is representable);
//   r^2 = -2 ln u:  lg2.approx has RELATIVE error 2^-22 outside (0.5, 2) but ABSOLUTE error 2^-22 inside, which is fatal
//        for r = sqrt(r^2) when u -> 1 (r -> 0: the two outliers of the round-1 SFU attempt).  For d = 1 - u < 2^-6 the
//        series -2 ln(1 - d) = 2 (d + d^2/2 + d^3/3) is used instead (truncation d^4/2 < 3e-8 -> |delta r| < 9e-8);
//        at d = 2^-6 the lg2 path has |delta r| = 2^-22 (2 ln 2) / (2 r) = 1.65e-7 / 0.177 = 9.3e-7, smaller beyond;
//   theta = 2 pi (u_b - 1/2) in (-pi, pi) (sin.approx / cos.approx: abs error 2^-20.9 there): t = f_b - 1.5 is exact,
//        theta = fma(t, 2 pi, 2 pi 2^-24) has one rounding; n = -r (cos, sin)(theta) because theta is shifted by pi.
__device__ __forceinline__ void box_muller_fast(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float fa = __uint_as_float((a >> 9) | 0x3F800000u);
  const float fb = __uint_as_float((b >> 9) | 0x3F800000u);
  const float u = fa - 0.99999994039535522f;                 // f - (1 - 2^-24), exact
  const float d = (2.0f - fa) - 5.9604644775390625e-08f;     // 1 - u = (2 - f) - 2^-24, exact
  float l2;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u));
  float r2 = l2 * -1.3862943611198906f;                      // -2 ln 2 log2 u
  const float ser = d * fmaf(d, fmaf(d, 0.66666666666666667f, 1.0f), 2.0f);   // 2 (d + d^2/2 + d^3/3)
  r2 = (d < 0.015625f) ? ser : r2;
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(r2));
  const float th = fmaf(fb - 1.5f, 6.2831853071795865f, 3.7450703e-07f);
  float sn, cs;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(th));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(th));
  n0 = -r * cs;
  n1 = -r * sn;
}
template <bool kExact>
__device__ __forceinline__ float4 normal4(uint64_t quad, uint32_t stream, uint32_t k0, uint32_t k1) {
  const uint4 w = philox4x32_10(make_uint4(static_cast<uint32_t>(quad), stream, static_cast<uint32_t>(quad >> 32), 0u),
                                k0, k1);
  float4 n;
  if (kExact) {
    box_muller(w.x, w.y, n.x, n.y);
    box_muller(w.z, w.w, n.z, n.w);
  } else {
    box_muller_fast(w.x, w.y, n.x, n.y);
    box_muller_fast(w.z, w.w, n.z, n.w);
  }
  return n;
}

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < kThreads / 32) ? sh[threadIdx.x] : 0.f;
  if (warp == 0) {
    t = warp_sum(t);
    if (lane == 0) sh[0] = t;
  }
  __syncthreads();
  const float r = sh[0];
  __syncthreads();
  return r;
}

struct ChanArgs {
  const float* z; int64_t z_sb, z_sm;
  float* y_f32; bf16* y_bf16; int64_t y_sb, y_sm;
  float* scale;
  int64_t B; int k; int kind; float rician_k;
  const float* snr_db; float snr_db_scalar;
  uint32_t key0, key1; int64_t sample_offset;
  const int64_t* sample_idx;                                // optional: global sample index of row b (else sample_offset + b)
  const uint64_t* seed_dev;                                 // optional device-side addend to the seed (CUDA-graph replays)
  const float* noise; const float* fading;
  float eps;
  int qlevels; float qclip;                                 // b-bit uniform quantiser of the normalised symbols (0 = off)
};

// Philox key of this launch: the by-value seed plus an optional device-side counter, so a captured graph draws fresh
// noise on every replay (the counter is bumped by a kernel inside the graph)
template <typename Args>
__device__ __forceinline__ uint2 chan_key(const Args& a) {
  uint64_t seed = (static_cast<uint64_t>(a.key1) << 32) | a.key0;
  if (a.seed_dev) seed += *a.seed_dev;
  return make_uint2(static_cast<uint32_t>(seed & 0xffffffffu), static_cast<uint32_t>(seed >> 32));
}
template <typename Args>
__device__ __forceinline__ int64_t global_sample(const Args& a, int64_t b) {
  return a.sample_idx ? a.sample_idx[b] : a.sample_offset + b;
}
// mid-rise uniform quantiser with L = 2^bits levels on [-clip, clip]: q = (clamp(floor(x / delta), -L/2, L/2 - 1) + 0.5) delta,
// delta = 2 clip / L (oracle/channel.py: quantize_ste).  Backward is the clipped straight-through estimator: 1 inside
// [-clip, clip], 0 outside.
__device__ __forceinline__ float quantize(float x, float inv_delta, float delta, float half_l) {
  const float i = fminf(fmaxf(floorf(x * inv_delta), -half_l), half_l - 1.0f);
  return (i + 0.5f) * delta;
}
__device__ __forceinline__ float4 quantize4(float4 s, int levels, float clip) {
  const float delta = 2.0f * clip / static_cast<float>(levels), inv_delta = static_cast<float>(levels) / (2.0f * clip);
  const float hl = 0.5f * static_cast<float>(levels);
  return make_float4(quantize(s.x, inv_delta, delta, hl), quantize(s.y, inv_delta, delta, hl),
                     quantize(s.z, inv_delta, delta, hl), quantize(s.w, inv_delta, delta, hl));
}
__device__ __forceinline__ float4 ste_mask4(float4 g, float4 zn, float clip) {
  return make_float4(fabsf(zn.x) <= clip ? g.x : 0.f, fabsf(zn.y) <= clip ? g.y : 0.f,
                     fabsf(zn.z) <= clip ? g.z : 0.f, fabsf(zn.w) <= clip ? g.w : 0.f);
}
__device__ __forceinline__ float sigma_of(const float* snr_db, float snr_scalar, int64_t b) {
  const float s = snr_db ? snr_db[b] : snr_scalar;
  return exp10f(-s * 0.05f);
}
// channel coefficient of (sample b, modality m): explicit tensor, or CN(0,1) from Philox stream 1 (Rayleigh), or
// sqrt(K/(K+1)) + sqrt(1/(K+1)) CN(0,1) (Rician, line of sight on the real axis; K = 0 is Rayleigh)
__device__ __forceinline__ float2 fading_of(const float* fading, int64_t b, int m, int64_t gb,
                                            uint32_t k0, uint32_t k1, int kind, float rician_k) {
  if (fading) return make_float2(fading[(b * 2 + m) * 2], fading[(b * 2 + m) * 2 + 1]);
  const uint64_t j = static_cast<uint64_t>(gb) * 2 + m;
  const uint4 w = philox4x32_10(make_uint4(static_cast<uint32_t>(j), 1u, static_cast<uint32_t>(j >> 32), 0u), k0, k1);
  float hr, hi;
  box_muller(w.x, w.y, hr, hi);
  hr *= 0.70710678118654752f;
  hi *= 0.70710678118654752f;
  if (kind == SC_CH_RICIAN) {
    const float los = sqrtf(rician_k / (rician_k + 1.0f)), sc = sqrtf(1.0f / (rician_k + 1.0f));
    hr = los + sc * hr;
    hi = sc * hi;
  }
  return make_float2(hr, hi);
}

template <bool kExact>
__global__ void __launch_bounds__(kThreads)
pnorm_channel_fwd_kernel(const ChanArgs a) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  const uint2 key = chan_key(a);
  __shared__ float sh[kThreads / 32];
  const int nv = a.k / 4;                                    // float4 per row
  for (int64_t row = blockIdx.x; row < a.B * 2; row += gridDim.x) {
    const int64_t b = row >> 1;
    const int m = static_cast<int>(row & 1);
    const float* zr = a.z + b * a.z_sb + m * a.z_sm;
    float4 v[kMaxV];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) {
      const int q = threadIdx.x + i * kThreads;
      if (q < nv) {
        v[i] = *reinterpret_cast<const float4*>(zr + q * 4);
        ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
      }
    }
    ss = block_sum(ss, sh);
    const float scale = rsqrtf(ss / static_cast<float>(a.k) + a.eps);
    if (threadIdx.x == 0 && a.scale) a.scale[row] = scale;
    float sigma = 0.f, hr = 1.f, hi = 0.f, inv_den = 1.f;
    if (a.kind != SC_CH_NONE) sigma = sigma_of(a.snr_db, a.snr_db_scalar, b);
    if (a.kind >= SC_CH_RAYLEIGH) {
      const float2 h = fading_of(a.fading, b, m, global_sample(a, b), key.x, key.y, a.kind, a.rician_k);
      hr = h.x; hi = h.y;
      inv_den = 1.0f / (hr * hr + hi * hi + sigma * sigma);
    }
    const uint64_t quad0 = (static_cast<uint64_t>(global_sample(a, b)) * 2 + m) * static_cast<uint64_t>(nv);
    float* yf = a.y_f32 ? a.y_f32 + b * a.y_sb + m * a.y_sm : nullptr;
    bf16* yb = a.y_bf16 ? a.y_bf16 + b * a.y_sb + m * a.y_sm : nullptr;
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) {
      const int q = threadIdx.x + i * kThreads;
      if (q < nv) {
        float4 s = make_float4(v[i].x * scale, v[i].y * scale, v[i].z * scale, v[i].w * scale);
        if (a.qlevels) s = quantize4(s, a.qlevels, a.qclip);
        float4 o = s;
        if (a.kind != SC_CH_NONE) {
          float4 n;
          if (a.noise) n = *reinterpret_cast<const float4*>(a.noise + row * a.k + q * 4);
          else n = normal4<kExact>(quad0 + q, 0u, key.x, key.y);
          if (a.kind == SC_CH_AWGN) {
            o = make_float4(fmaf(sigma, n.x, s.x), fmaf(sigma, n.y, s.y), fmaf(sigma, n.z, s.z), fmaf(sigma, n.w, s.w));
          } else {
            // two complex symbols (x,y) and (z,w): r = h s + sigma n ; y = conj(h) r / (|h|^2 + sigma^2)
            const float r0r = hr * s.x - hi * s.y + sigma * n.x, r0i = hr * s.y + hi * s.x + sigma * n.y;
            const float r1r = hr * s.z - hi * s.w + sigma * n.z, r1i = hr * s.w + hi * s.z + sigma * n.w;
            o = make_float4((hr * r0r + hi * r0i) * inv_den, (hr * r0i - hi * r0r) * inv_den,
                            (hr * r1r + hi * r1i) * inv_den, (hr * r1i - hi * r1r) * inv_den);
          }
        }
        if (yf) *reinterpret_cast<float4*>(yf + q * 4) = o;
        if (yb) *reinterpret_cast<uint2*>(yb + q * 4) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      }
    }
  }
}

struct ChanBwdArgs {
  const float* dy; int64_t y_sb, y_sm;
  const float* z; int64_t z_sb, z_sm;
  const float* scale;
  float* dz_f32; bf16* dz_bf16;
  int64_t B; int k; int kind; float rician_k;
  const float* snr_db; float snr_db_scalar;
  uint32_t key0, key1; int64_t sample_offset;
  const int64_t* sample_idx;
  const uint64_t* seed_dev;
  const float* fading;
  int qlevels; float qclip;
};

__global__ void __launch_bounds__(kThreads)
pnorm_channel_bwd_kernel(const ChanBwdArgs a) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  const uint2 key = chan_key(a);
  __shared__ float sh[kThreads / 32];
  const int nv = a.k / 4;
  for (int64_t row = blockIdx.x; row < a.B * 2; row += gridDim.x) {
    const int64_t b = row >> 1;
    const int m = static_cast<int>(row & 1);
    const float* zr = a.z + b * a.z_sb + m * a.z_sm;
    const float* gr = a.dy + b * a.y_sb + m * a.y_sm;
    const float scale = a.scale[row];
    float gain = 1.f;
    if (a.kind >= SC_CH_RAYLEIGH) {
      const float sigma = sigma_of(a.snr_db, a.snr_db_scalar, b);
      const float2 h = fading_of(a.fading, b, m, global_sample(a, b), key.x, key.y, a.kind, a.rician_k);
      const float h2 = h.x * h.x + h.y * h.y;
      gain = h2 / (h2 + sigma * sigma);
    }
    float4 zn[kMaxV], g[kMaxV];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) {
      const int q = threadIdx.x + i * kThreads;
      if (q < nv) {
        const float4 zv = *reinterpret_cast<const float4*>(zr + q * 4);
        const float4 gv = *reinterpret_cast<const float4*>(gr + q * 4);
        zn[i] = make_float4(zv.x * scale, zv.y * scale, zv.z * scale, zv.w * scale);
        g[i] = make_float4(gv.x * gain, gv.y * gain, gv.z * gain, gv.w * gain);
        if (a.qlevels) g[i] = ste_mask4(g[i], zn[i], a.qclip);
        dot += g[i].x * zn[i].x + g[i].y * zn[i].y + g[i].z * zn[i].z + g[i].w * zn[i].w;
      }
    }
    dot = block_sum(dot, sh) / static_cast<float>(a.k);
    float* df = a.dz_f32 ? a.dz_f32 + b * a.z_sb + m * a.z_sm : nullptr;
    bf16* db = a.dz_bf16 ? a.dz_bf16 + b * a.z_sb + m * a.z_sm : nullptr;
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) {
      const int q = threadIdx.x + i * kThreads;
      if (q < nv) {
        const float4 o = make_float4(scale * (g[i].x - zn[i].x * dot), scale * (g[i].y - zn[i].y * dot),
                                     scale * (g[i].z - zn[i].z * dot), scale * (g[i].w - zn[i].w * dot));
        if (df) *reinterpret_cast<float4*>(df + q * 4) = o;
        if (db) *reinterpret_cast<uint2*>(db + q * 4) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      }
    }
  }
}

__global__ void philox_words_kernel(uint32_t* out, int64_t n_quads, uint32_t stream_id, uint32_t k0, uint32_t k1,
                                    int64_t quad_offset) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_quads;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint64_t q = static_cast<uint64_t>(quad_offset + i);
    const uint4 w = philox4x32_10(make_uint4(static_cast<uint32_t>(q), stream_id, static_cast<uint32_t>(q >> 32), 0u), k0, k1);
    *reinterpret_cast<uint4*>(out + i * 4) = w;
  }
}

// ------------------------------------------------------------------------------------------
// warp-per-row variants (k <= 2048): the row lives in one warp's registers, the reduction is pure
// shuffles, and 8 independent rows are in flight per CTA — no block barrier on the critical path.
// ------------------------------------------------------------------------------------------
constexpr int kWarpV = 16;                    // float4 per lane -> k <= 2048

template <int NV, bool kExact>
__global__ void __launch_bounds__(kThreads)
pnorm_channel_fwd_warp_kernel(const ChanArgs a) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  const uint2 key = chan_key(a);
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int nv = a.k / 4;
  for (int64_t row = warp0; row < a.B * 2; row += nwarps) {
    const int64_t b = row >> 1;
    const int m = static_cast<int>(row & 1);
    const float* zr = a.z + b * a.z_sb + m * a.z_sm;
    float4 v[NV];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int q = lane + i * 32;
      if (q < nv) {
        v[i] = *reinterpret_cast<const float4*>(zr + q * 4);
        ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
      }
    }
    ss = warp_sum(ss);
    const float scale = rsqrtf(ss / static_cast<float>(a.k) + a.eps);
    if (lane == 0 && a.scale) a.scale[row] = scale;
    float sigma = 0.f, hr = 1.f, hi = 0.f, inv_den = 1.f;
    if (a.kind != SC_CH_NONE) sigma = sigma_of(a.snr_db, a.snr_db_scalar, b);
    if (a.kind >= SC_CH_RAYLEIGH) {
      const float2 h = fading_of(a.fading, b, m, global_sample(a, b), key.x, key.y, a.kind, a.rician_k);
      hr = h.x; hi = h.y;
      inv_den = 1.0f / (hr * hr + hi * hi + sigma * sigma);
    }
    const uint64_t quad0 = (static_cast<uint64_t>(global_sample(a, b)) * 2 + m) * static_cast<uint64_t>(nv);
    float* yf = a.y_f32 ? a.y_f32 + b * a.y_sb + m * a.y_sm : nullptr;
    bf16* yb = a.y_bf16 ? a.y_bf16 + b * a.y_sb + m * a.y_sm : nullptr;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int q = lane + i * 32;
      if (q < nv) {
        float4 s = make_float4(v[i].x * scale, v[i].y * scale, v[i].z * scale, v[i].w * scale);
        if (a.qlevels) s = quantize4(s, a.qlevels, a.qclip);
        float4 o = s;
        if (a.kind != SC_CH_NONE) {
          float4 n;
          if (a.noise) n = *reinterpret_cast<const float4*>(a.noise + row * a.k + q * 4);
          else n = normal4<kExact>(quad0 + q, 0u, key.x, key.y);
          if (a.kind == SC_CH_AWGN) {
            o = make_float4(fmaf(sigma, n.x, s.x), fmaf(sigma, n.y, s.y), fmaf(sigma, n.z, s.z), fmaf(sigma, n.w, s.w));
          } else {
            const float r0r = hr * s.x - hi * s.y + sigma * n.x, r0i = hr * s.y + hi * s.x + sigma * n.y;
            const float r1r = hr * s.z - hi * s.w + sigma * n.z, r1i = hr * s.w + hi * s.z + sigma * n.w;
            o = make_float4((hr * r0r + hi * r0i) * inv_den, (hr * r0i - hi * r0r) * inv_den,
                            (hr * r1r + hi * r1i) * inv_den, (hr * r1i - hi * r1r) * inv_den);
          }
        }
        if (yf) *reinterpret_cast<float4*>(yf + q * 4) = o;
        if (yb) *reinterpret_cast<uint2*>(yb + q * 4) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// The production shape (k = 1024 reals per row, in-kernel noise, continuous symbols) as a compile-time specialisation:
// the channel kind is a template parameter (no branches in the symbol loop), all 8 float4 of a lane are live (no
// predicates), the ten Philox round keys are hoisted out of the row loop, and the fading channel is applied in the
// folded form  y = (|h|^2 s + sigma conj(h) n) / (|h|^2 + sigma^2)  (three operations per real symbol).
// ncu (round 1) had the generic kernel issue-bound at 64 thread-instructions per symbol; this one needs ~26.
// ------------------------------------------------------------------------------------------
struct PhiloxKeys {
  uint32_t k0[10], k1[10];
};
__device__ __forceinline__ PhiloxKeys philox_keys(uint32_t k0, uint32_t k1) {
  PhiloxKeys ks;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    ks.k0[r] = k0 + static_cast<uint32_t>(r) * PHILOX_W0;
    ks.k1[r] = k1 + static_cast<uint32_t>(r) * PHILOX_W1;
  }
  return ks;
}
__device__ __forceinline__ uint4 philox4x32_10_keys(uint4 c, const PhiloxKeys& ks) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(PHILOX_M0) * c.x;
    const uint64_t p1 = static_cast<uint64_t>(PHILOX_M1) * c.z;
    c = make_uint4(static_cast<uint32_t>(p1 >> 32) ^ c.y ^ ks.k0[r], static_cast<uint32_t>(p1),
                   static_cast<uint32_t>(p0 >> 32) ^ c.w ^ ks.k1[r], static_cast<uint32_t>(p0));
  }
  return c;
}

template <int KIND, bool kExact>
__global__ void __launch_bounds__(kThreads)
pnorm_channel_fwd_k1024_kernel(const ChanArgs a) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  constexpr int NV = 8;                                      // 8 float4 per lane = 1024 reals per warp
  const uint2 key = chan_key(a);
  const PhiloxKeys ks = philox_keys(key.x, key.y);
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t row = warp0; row < a.B * 2; row += nwarps) {
    const int64_t b = row >> 1;
    const int m = static_cast<int>(row & 1);
    const float4* zr = reinterpret_cast<const float4*>(a.z + b * a.z_sb + m * a.z_sm) + lane;
    float4 v[NV];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = zr[i * 32];
#pragma unroll
    for (int i = 0; i < NV; ++i) ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    ss = warp_sum(ss);
    const float scale = rsqrtf(ss * (1.0f / 1024.0f) + a.eps);
    if (lane == 0 && a.scale) a.scale[row] = scale;
    const int64_t gb = global_sample(a, b);
    // symbol gain and noise coefficients: AWGN y = s + sigma n; fading y = g s + (ar n_r + ai n_i, ar n_i - ai n_r)
    float g = scale, ar = 0.f, ai = 0.f;
    if (KIND != SC_CH_NONE) {
      const float sigma = sigma_of(a.snr_db, a.snr_db_scalar, b);
      ar = sigma;
      if (KIND >= SC_CH_RAYLEIGH) {
        const float2 h = fading_of(a.fading, b, m, gb, key.x, key.y, KIND, a.rician_k);
        const float h2 = h.x * h.x + h.y * h.y;
        const float inv_den = 1.0f / (h2 + sigma * sigma);
        g = scale * h2 * inv_den;
        ar = sigma * h.x * inv_den;
        ai = sigma * h.y * inv_den;
      }
    }
    const uint64_t quad0 = (static_cast<uint64_t>(gb) * 2 + m) * 256u + static_cast<uint32_t>(lane);
    float4* yf = a.y_f32 ? reinterpret_cast<float4*>(a.y_f32 + b * a.y_sb + m * a.y_sm) + lane : nullptr;
    uint2* yb = a.y_bf16 ? reinterpret_cast<uint2*>(a.y_bf16 + b * a.y_sb + m * a.y_sm) + lane : nullptr;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o = make_float4(v[i].x * g, v[i].y * g, v[i].z * g, v[i].w * g);
      if (KIND != SC_CH_NONE) {
        const uint64_t quad = quad0 + static_cast<uint32_t>(i * 32);
        const uint4 w = philox4x32_10_keys(make_uint4(static_cast<uint32_t>(quad), 0u, static_cast<uint32_t>(quad >> 32), 0u), ks);
        float4 n;
        if (kExact) {
          box_muller(w.x, w.y, n.x, n.y);
          box_muller(w.z, w.w, n.z, n.w);
        } else {
          box_muller_fast(w.x, w.y, n.x, n.y);
          box_muller_fast(w.z, w.w, n.z, n.w);
        }
        if (KIND == SC_CH_AWGN) {
          o = make_float4(fmaf(ar, n.x, o.x), fmaf(ar, n.y, o.y), fmaf(ar, n.z, o.z), fmaf(ar, n.w, o.w));
        } else {
          o = make_float4(fmaf(ai, n.y, fmaf(ar, n.x, o.x)), fmaf(-ai, n.x, fmaf(ar, n.y, o.y)),
                          fmaf(ai, n.w, fmaf(ar, n.z, o.z)), fmaf(-ai, n.z, fmaf(ar, n.w, o.w)));
        }
      }
      if (yf) yf[i * 32] = o;
      if (yb) yb[i * 32] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(kThreads)
pnorm_channel_bwd_warp_kernel(const ChanBwdArgs a) {
  pdl_sync();                                   // first statement: no global-memory access before it (common.cuh)
  const uint2 key = chan_key(a);
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int nv = a.k / 4;
  for (int64_t row = warp0; row < a.B * 2; row += nwarps) {
    const int64_t b = row >> 1;
    const int m = static_cast<int>(row & 1);
    const float* zr = a.z + b * a.z_sb + m * a.z_sm;
    const float* gr = a.dy + b * a.y_sb + m * a.y_sm;
    const float scale = a.scale[row];
    float gain = 1.f;
    if (a.kind >= SC_CH_RAYLEIGH) {
      const float sigma = sigma_of(a.snr_db, a.snr_db_scalar, b);
      const float2 h = fading_of(a.fading, b, m, global_sample(a, b), key.x, key.y, a.kind, a.rician_k);
      const float h2 = h.x * h.x + h.y * h.y;
      gain = h2 / (h2 + sigma * sigma);
    }
    float4 zn[NV], g[NV];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int q = lane + i * 32;
      if (q < nv) {
        const float4 zv = *reinterpret_cast<const float4*>(zr + q * 4);
        const float4 gv = *reinterpret_cast<const float4*>(gr + q * 4);
        zn[i] = make_float4(zv.x * scale, zv.y * scale, zv.z * scale, zv.w * scale);
        g[i] = make_float4(gv.x * gain, gv.y * gain, gv.z * gain, gv.w * gain);
        if (a.qlevels) g[i] = ste_mask4(g[i], zn[i], a.qclip);
        dot += g[i].x * zn[i].x + g[i].y * zn[i].y + g[i].z * zn[i].z + g[i].w * zn[i].w;
      }
    }
    dot = warp_sum(dot) / static_cast<float>(a.k);
    float* df = a.dz_f32 ? a.dz_f32 + b * a.z_sb + m * a.z_sm : nullptr;
    bf16* db = a.dz_bf16 ? a.dz_bf16 + b * a.z_sb + m * a.z_sm : nullptr;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int q = lane + i * 32;
      if (q < nv) {
        const float4 o = make_float4(scale * (g[i].x - zn[i].x * dot), scale * (g[i].y - zn[i].y * dot),
                                     scale * (g[i].z - zn[i].z * dot), scale * (g[i].w - zn[i].w * dot));
        if (df) *reinterpret_cast<float4*>(df + q * 4) = o;
        if (db) *reinterpret_cast<uint2*>(db + q * 4) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      }
    }
  }
}

int check_common(const char* nm, int64_t B, int k, int kind, int64_t sb, int64_t sm) {
  SC_REQUIRE(B > 0 && k > 0 && k % 4 == 0 && k <= kMaxV * kThreads * 4, SC_E_SHAPE,
             "%s: B=%lld k=%d (need k %% 4 == 0 and k <= %d)", nm, (long long)B, k, kMaxV * kThreads * 4);
  SC_REQUIRE(kind >= SC_CH_NONE && kind <= SC_CH_RICIAN, SC_E_SHAPE, "%s: unknown channel kind %d", nm, kind);
  SC_REQUIRE(sb % 4 == 0 && sm % 4 == 0, SC_E_ALIGN, "%s: strides must be multiples of 4 elements", nm);
  return SC_OK;
}

}  // namespace

extern "C" {

int g_chan_exact = 0;          // sc_debug_set("chan_exact", 1): accurate logf / sincospif Box-Muller (A/B yardstick)
int g_chan_generic = 0;        // sc_debug_set("chan_generic", 1): the run-time-kind kernel also for the production shape

int sc_pnorm_channel_fwd(const float* z, int64_t z_sb, int64_t z_sm, float* y_f32, void* y_bf16, int64_t y_sb,
                         int64_t y_sm, float* scale, int64_t B, int k, int kind, float rician_k,
                         const float* snr_db, float snr_db_scalar, uint64_t seed, int64_t sample_offset,
                         const int64_t* sample_idx, const float* noise, const float* fading, float eps,
                         const uint64_t* seed_dev, int quant_bits, float quant_clip, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(z && (y_f32 || y_bf16), SC_E_SHAPE, "sc_pnorm_channel_fwd: null pointer");
  if ((rc = check_common("sc_pnorm_channel_fwd", B, k, kind, z_sb, z_sm)) != SC_OK) return rc;
  SC_REQUIRE(kind != SC_CH_RICIAN || rician_k >= 0.f, SC_E_SHAPE, "sc_pnorm_channel_fwd: rician_k must be >= 0");
  SC_REQUIRE(quant_bits >= 0 && quant_bits <= 16 && (quant_bits == 0 || quant_clip > 0.f), SC_E_SHAPE,
             "sc_pnorm_channel_fwd: quant_bits must be in [0, 16] and quant_clip > 0");
  if ((rc = check_common("sc_pnorm_channel_fwd", B, k, kind, y_sb, y_sm)) != SC_OK) return rc;
  SC_REQUIRE(sc_aligned16(z) && sc_aligned16(y_f32) && sc_aligned16(noise) &&
                 (reinterpret_cast<uintptr_t>(y_bf16) & 7) == 0,
             SC_E_ALIGN, "sc_pnorm_channel_fwd: unaligned pointer");
  ChanArgs a;
  a.z = z; a.z_sb = z_sb; a.z_sm = z_sm;
  a.y_f32 = y_f32; a.y_bf16 = (bf16*)y_bf16; a.y_sb = y_sb; a.y_sm = y_sm;
  a.scale = scale; a.B = B; a.k = k; a.kind = kind; a.rician_k = rician_k;
  a.snr_db = snr_db; a.snr_db_scalar = snr_db_scalar;
  a.key0 = (uint32_t)(seed & 0xffffffffu); a.key1 = (uint32_t)(seed >> 32);
  a.sample_offset = sample_offset; a.sample_idx = sample_idx; a.noise = noise; a.fading = fading; a.eps = eps;
  a.seed_dev = seed_dev;
  a.qlevels = quant_bits > 0 ? (1 << quant_bits) : 0; a.qclip = quant_clip;
  const int64_t rows = B * 2;
  const bool exact = g_chan_exact != 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (k == 1024 && !noise && !a.qlevels && g_chan_generic == 0) {        // the production shape: specialised kernels
    const int64_t blocks = (rows + 7) / 8, cap = (int64_t)sc_num_sms() * 8;
    const dim3 grid((int)(blocks < cap ? blocks : cap));
#define SC_CHAN_LAUNCH(KIND)                                                                                         \
  do {                                                                                                               \
    if (exact) sc_launch_kernel(pnorm_channel_fwd_k1024_kernel<KIND, true>, grid, dim3(kThreads), 0, st, a);         \
    else sc_launch_kernel(pnorm_channel_fwd_k1024_kernel<KIND, false>, grid, dim3(kThreads), 0, st, a);              \
  } while (0)
    switch (kind) {
      case SC_CH_NONE: SC_CHAN_LAUNCH(SC_CH_NONE); break;
      case SC_CH_AWGN: SC_CHAN_LAUNCH(SC_CH_AWGN); break;
      case SC_CH_RAYLEIGH: SC_CHAN_LAUNCH(SC_CH_RAYLEIGH); break;
      default: SC_CHAN_LAUNCH(SC_CH_RICIAN); break;
    }
#undef SC_CHAN_LAUNCH
  } else if (k <= 1024) {                        // warp per row, 8 rows per CTA
    const int64_t blocks = (rows + 7) / 8, cap = (int64_t)sc_num_sms() * 8;
    const dim3 grid((int)(blocks < cap ? blocks : cap));
    if (exact) sc_launch_kernel(pnorm_channel_fwd_warp_kernel<8, true>, grid, dim3(kThreads), 0, st, a);
    else sc_launch_kernel(pnorm_channel_fwd_warp_kernel<8, false>, grid, dim3(kThreads), 0, st, a);
  } else if (k <= kWarpV * 128) {
    const int64_t blocks = (rows + 7) / 8, cap = (int64_t)sc_num_sms() * 4;
    const dim3 grid((int)(blocks < cap ? blocks : cap));
    if (exact) sc_launch_kernel(pnorm_channel_fwd_warp_kernel<kWarpV, true>, grid, dim3(kThreads), 0, st, a);
    else sc_launch_kernel(pnorm_channel_fwd_warp_kernel<kWarpV, false>, grid, dim3(kThreads), 0, st, a);
  } else {
    const int64_t cap = (int64_t)sc_num_sms() * 8;
    const dim3 grid((int)(rows < cap ? rows : cap));
    if (exact) sc_launch_kernel(pnorm_channel_fwd_kernel<true>, grid, dim3(kThreads), 0, st, a);
    else sc_launch_kernel(pnorm_channel_fwd_kernel<false>, grid, dim3(kThreads), 0, st, a);
  }
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_pnorm_channel_fwd");
  return SC_OK;
}

int sc_pnorm_channel_bwd(const float* dy, int64_t y_sb, int64_t y_sm, const float* z, int64_t z_sb, int64_t z_sm,
                         const float* scale, float* dz_f32, void* dz_bf16, int64_t B, int k, int kind, float rician_k,
                         const float* snr_db, float snr_db_scalar, uint64_t seed, int64_t sample_offset,
                         const int64_t* sample_idx, const float* fading, const uint64_t* seed_dev, int quant_bits,
                         float quant_clip, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(dy && z && scale && (dz_f32 || dz_bf16), SC_E_SHAPE, "sc_pnorm_channel_bwd: null pointer");
  if ((rc = check_common("sc_pnorm_channel_bwd", B, k, kind, z_sb, z_sm)) != SC_OK) return rc;
  SC_REQUIRE(kind != SC_CH_RICIAN || rician_k >= 0.f, SC_E_SHAPE, "sc_pnorm_channel_bwd: rician_k must be >= 0");
  SC_REQUIRE(quant_bits >= 0 && quant_bits <= 16 && (quant_bits == 0 || quant_clip > 0.f), SC_E_SHAPE,
             "sc_pnorm_channel_bwd: quant_bits must be in [0, 16] and quant_clip > 0");
  if ((rc = check_common("sc_pnorm_channel_bwd", B, k, kind, y_sb, y_sm)) != SC_OK) return rc;
  SC_REQUIRE(sc_aligned16(dy) && sc_aligned16(z) && sc_aligned16(dz_f32) &&
                 (reinterpret_cast<uintptr_t>(dz_bf16) & 7) == 0,
             SC_E_ALIGN, "sc_pnorm_channel_bwd: unaligned pointer");
  ChanBwdArgs a;
  a.dy = dy; a.y_sb = y_sb; a.y_sm = y_sm; a.z = z; a.z_sb = z_sb; a.z_sm = z_sm; a.scale = scale;
  a.dz_f32 = dz_f32; a.dz_bf16 = (bf16*)dz_bf16; a.B = B; a.k = k; a.kind = kind; a.rician_k = rician_k;
  a.snr_db = snr_db; a.snr_db_scalar = snr_db_scalar;
  a.key0 = (uint32_t)(seed & 0xffffffffu); a.key1 = (uint32_t)(seed >> 32);
  a.sample_offset = sample_offset; a.sample_idx = sample_idx; a.fading = fading; a.seed_dev = seed_dev;
  a.qlevels = quant_bits > 0 ? (1 << quant_bits) : 0; a.qclip = quant_clip;
  const int64_t rows = B * 2;
  if (k <= 1024) {
    const int64_t blocks = (rows + 7) / 8, cap = (int64_t)sc_num_sms() * 6;
    sc_launch_kernel(pnorm_channel_bwd_warp_kernel<8>, dim3((int)(blocks < cap ? blocks : cap)), dim3(kThreads), 0, (cudaStream_t)stream, a);
  } else if (k <= kWarpV * 128) {
    const int64_t blocks = (rows + 7) / 8, cap = (int64_t)sc_num_sms() * 3;
    sc_launch_kernel(pnorm_channel_bwd_warp_kernel<kWarpV>, dim3((int)(blocks < cap ? blocks : cap)), dim3(kThreads), 0, (cudaStream_t)stream, a);
  } else {
    const int64_t cap = (int64_t)sc_num_sms() * 8;
    sc_launch_kernel(pnorm_channel_bwd_kernel, dim3((int)(rows < cap ? rows : cap)), dim3(kThreads), 0, (cudaStream_t)stream, a);
  }
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_pnorm_channel_bwd");
  return SC_OK;
}

}  // extern "C"

namespace {
template <bool kExact>
__global__ void philox_normals_kernel(float* out, int64_t n_quads, uint32_t k0, uint32_t k1, int64_t quad_offset) {
  pdl_sync();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_quads;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 n = normal4<kExact>(static_cast<uint64_t>(quad_offset + i), 0u, k0, k1);
    *reinterpret_cast<float4*>(out + i * 4) = n;
  }
}
}  // namespace

extern "C" {

// raw normals of the in-kernel generator (parity / error-bound tests): out fp32 [n_quads, 4], stream 0 counters
// quad_offset .. ; exact != 0 selects the accurate Box-Muller
int sc_philox_normals(float* out, int64_t n_quads, uint64_t seed, int64_t quad_offset, int exact, void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(out && n_quads > 0, SC_E_SHAPE, "sc_philox_normals: bad arguments");
  SC_REQUIRE(sc_aligned16(out), SC_E_ALIGN, "sc_philox_normals: unaligned pointer");
  const int grid = (int)((n_quads + 255) / 256 < 1184 ? (n_quads + 255) / 256 : 1184);
  const uint32_t k0 = (uint32_t)(seed & 0xffffffffu), k1 = (uint32_t)(seed >> 32);
  if (exact) sc_launch_kernel(philox_normals_kernel<true>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, out, n_quads, k0, k1, quad_offset);
  else sc_launch_kernel(philox_normals_kernel<false>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, out, n_quads, k0, k1, quad_offset);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_philox_normals");
  return SC_OK;
}

int sc_philox_words(uint32_t* out, int64_t n_quads, uint32_t stream_id, uint64_t seed, int64_t quad_offset,
                    void* stream) {
  int rc = sc_check_device();
  if (rc != SC_OK) return rc;
  SC_REQUIRE(out && n_quads > 0, SC_E_SHAPE, "sc_philox_words: bad arguments");
  SC_REQUIRE(sc_aligned16(out), SC_E_ALIGN, "sc_philox_words: unaligned pointer");
  const int grid = (int)((n_quads + 255) / 256 < 1184 ? (n_quads + 255) / 256 : 1184);
  sc_launch_kernel(philox_words_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, out, n_quads, stream_id, (uint32_t)(seed & 0xffffffffu),
                                                              (uint32_t)(seed >> 32), quad_offset);
  sc_count_launch(1);
  SC_CHECK_LAUNCH("sc_philox_words");
  return SC_OK;
}

}  // extern "C"
