/* semcom_sm100.h — C ABI of libsemcom_sm100.so (hand-written sm_100a kernels).
 *
 * Drop-in boundary of the SemCom-PPO hot path.  The upstream repository
 * (/root/reference/README.md:1-2) contains no code and therefore no plugin / FFI interface
 * to replace; each entry point below cites the SPEC v0 row it implements
 * instead, and INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless noted;
 *   - the caller owns every buffer; the library never allocates, frees or retains pointers;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), no internal
 *     synchronisation, no default-stream use -> CUDA-graph capturable;
 *   - return 0 (SC_OK) or a negative SC_E_* code; sc_last_error() gives the thread-local
 *     message; no exceptions, no exit();
 *   - tensors are contiguous row-major unless a leading dimension is passed, 16-byte aligned;
 *   - a device that is not compute capability 10.x is refused with SC_E_ARCH (no fallback).
 */
#ifndef SEMCOM_SM100_H_
#define SEMCOM_SM100_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SC_OK 0
#define SC_E_SHAPE (-1)
#define SC_E_ALIGN (-2)
#define SC_E_ARCH (-3)
#define SC_E_LAUNCH (-4)
#define SC_E_DRIVER (-5)

#define SC_ABI_VERSION 4

/* channel kinds (SPEC v0 a4) */
#define SC_CH_NONE 0
#define SC_CH_AWGN 1
#define SC_CH_RAYLEIGH 2
#define SC_CH_RICIAN 3   /* block Rician fading: h = sqrt(K/(K+1)) + sqrt(1/(K+1)) CN(0,1), MMSE-equalised like Rayleigh */

/* activation ids for the GEMM epilogue */
#define SC_ACT_NONE 0
#define SC_ACT_GELU 1 /* tanh-form GELU */
#define SC_ACT_TANH 2
#define SC_ACT_GELU_ERF 3 /* exact GELU x Phi(x) (SURVEY.md 8 froze "GELU(tanh approx = False)"; SpecV0.gelu = "erf") */

int sc_version(void);
const char* sc_last_error(void);
/* SC_OK when the current device is sm_100; SC_E_ARCH otherwise. */
int sc_device_ok(void);
/* number of kernels launched by this library since load (all threads); bench.py's gpu_launches */
int64_t sc_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * Dense contraction on tcgen05 tensor cores, operands by TMA (SPEC v0 a1/a2/a5 GEMMs:
 * patch-embed, QKV, out-proj, MLP, symbol/in/recon projections, and their dgrad/wgrad).
 *
 *   acc[m,n] = sum_k A(m,k) * B(n,k)                 bf16 x bf16 -> fp32 (TMEM accumulators)
 *   a_mn = 0: A is stored [M,K] row-major (lda);  a_mn = 1: A is stored [K,M] row-major (lda)
 *   b_mn = 0: B is stored [N,K] row-major (ldb);  b_mn = 1: B is stored [K,N] row-major (ldb)
 *   so  fwd   y = x W^T      : (a_mn,b_mn) = (0,0)
 *       dgrad dx = dy W      : (0,1)   (W itself, no transposed copy)
 *       wgrad dW = dy^T x    : (1,1)   (reduction over the huge M, use ksplit > 1 + atomic)
 *
 * epilogue, applied per element in this order (null pointer / zero = skipped):
 *   v = alpha * acc
 *   v += bias[n]                                    fp32 [N]
 *   v += rowadd[(m % rowadd_period) * N + n]        fp32 [period, N]     (pos-embed; or rowadd_bf16)
 *   v += rowgrp[(m / rowgrp_size) * N + n] * rowgrp_scale   fp32          (mean-pool bwd)
 *   aux_out[m*ldo + n] = bf16(v)                    pre-activation copy   (GELU fwd)
 *   v = act(v)
 *   v *= gelu'(aux_in[m*ldo + n])                   bf16                  (GELU bwd)
 *   v += resid[m*ldo + n]                           bf16                  (residual)
 *   out_bf16[m*ldo + n] = bf16(v)  and/or  out_f32[m*ldo + n] (= or +=) v
 * ldo is the leading dimension of out_bf16 / out_f32 / aux_out / aux_in / resid.
 * atomic != 0 accumulates out_f32 with red.global.add (required when ksplit > 1).
 * Requirements: lda, ldb multiples of 8; A, B 16-byte aligned; K > 0.
 * ------------------------------------------------------------------------------------- */
typedef struct sc_gemm_epilogue {
  const float* bias;
  const float* rowadd;
  int64_t rowadd_period;
  const float* rowgrp;
  int64_t rowgrp_size;
  float rowgrp_scale;
  float alpha;
  const void* resid;  /* bf16 */
  const void* aux_in; /* bf16 */
  void* aux_out;      /* bf16 */
  void* out_bf16;
  float* out_f32;
  int64_t ldo;
  int32_t act;
  int32_t atomic;
  const void* rowadd_bf16; /* bf16 [rowadd_period, N]: same add as `rowadd`, but fetched as a TMA tile
                              (needs bf16 output, N % 256 == 0, period % 32 == 0, no resid / aux_in) */
  float* colsum;           /* a_mn = 1 (wgrad form) only: colsum[m] += sum_k A(m,k), i.e. the bias gradient
                              sum_tokens dy[token, m], computed by one extra tcgen05.mma per K-slice against
                              an all-ones shared-memory tile (no extra memory traffic) */
  /* fused LayerNorm of the output row (N == 256, bf16 output with bias and one bf16 input tile): besides
   * out_bf16 = acc + bias + resid the kernel writes ln_out = LayerNorm(out_bf16) * ln_gamma + ln_beta (bf16,
   * same leading dimension) and the row statistics (fp32 [M], optional) that sc_layernorm_bwd consumes. */
  void* ln_out;
  const float* ln_gamma;
  const float* ln_beta;
  float* ln_mean;
  float* ln_rstd;
  float ln_eps;
  int32_t aux_in_erf;      /* != 0: the GELU' applied to aux_in is the exact (erf) form instead of the tanh form */
} sc_gemm_epilogue;

int sc_gemm_bf16(const void* A, const void* B, int64_t M, int64_t N, int64_t K, int a_mn, int b_mn,
                 int64_t lda, int64_t ldb, const sc_gemm_epilogue* epi, int ksplit, void* stream);

/* ---------------------------------------------------------------------------------------
 * Transformer MLP as one kernel per direction (SPEC v0 a1/a2/a5: the block's fc1 -> GELU -> fc2).
 * The 128 x R hidden tile is produced chunk by chunk in TMEM, transformed by the epilogue warps and consumed
 * by the second GEMM from shared memory; it never makes an HBM round trip between the two GEMMs.
 *   sc_mlp_fwd:        out = resid + gelu(x W1^T + b1) W2^T + b2;   hpre (pre-activation) and h (activation)
 *                      are saved for backward when the pointers are non-null (both bf16 [M, R]).
 *   sc_mlp_bwd_dgrad:  dhpre = (dy W2) o gelu'(hpre)  (saved, bf16 [M, R]: the fc1 wgrad operand),
 *                      dx = dhpre W1.   The two weight gradients stay sc_gemm_bf16 calls.
 * x/dy/out/dx/resid: bf16 [M, D] row-major, D must be 256; W1: bf16 [R, D]; W2: bf16 [D, R]; R % 128 == 0.
 * ------------------------------------------------------------------------------------- */
int sc_mlp_fwd(const void* x, const void* w1, const float* b1, const void* w2, const float* b2,
               const void* resid, void* hpre, void* h, void* out, int64_t M, int D, int R, void* stream);
int sc_mlp_bwd_dgrad(const void* dy, const void* w1, const void* w2, const void* hpre, void* dhpre, void* dx,
                     int64_t M, int D, int R, void* stream);
/* inference form of sc_mlp_fwd (no hpre / h saves) with the LayerNorm that consumes the block's output fused into the
 * output epilogue: out = resid + W2 gelu(W1 x + b1) + b2 and ln_out = LayerNorm(out) * gamma + beta (statistics of the
 * bf16-rounded out, like sc_layernorm_fwd on the stored tensor; ln_mean / ln_rstd fp32 [M], optional) — one launch less per
 * transformer block where launches, not bytes, set the pace (cfg2, act()). */
int sc_mlp_fwd_ln(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, const void* resid,
                  void* out, const float* ln_gamma, const float* ln_beta, float ln_eps, void* ln_out, float* ln_mean,
                  float* ln_rstd, int64_t M, int D, int R, void* stream);

/* test hook: override UMMA descriptor fields (0 = library default). Not for product use. */
int sc_debug_set(const char* key, int64_t value);

/* ---------------------------------------------------------------------------------------
 * Fused multi-head self-attention for S in {64,128}, d_head = 64 (SPEC v0 a1/a2/a5 blocks).
 * qkv: bf16 [B*S, 3*H*64], columns [q(h,d) | k(h,d) | v(h,d)];  out: bf16 [B*S, H*64];
 * lse: fp32 [B*S, H] (log2-domain log-sum-exp of the scaled scores, saved for backward).
 * Whole S x S score tile lives in TMEM; QK^T and PV on tcgen05.
 * ------------------------------------------------------------------------------------- */
int sc_attn_fwd(const void* qkv, void* out, float* lse, int64_t B, int S, int H, void* stream);
int sc_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                int64_t B, int S, int H, void* stream);

/* ---------------------------------------------------------------------------------------
 * Memory-bound fused kernels.
 * ------------------------------------------------------------------------------------- */
/* a1/a2: img [*,C,H,W] (fp32, or bf16 when img_is_bf16 != 0) -> patch tokens bf16
 * [B*(H/P)*(W/P), C*P*P], feature order (c,py,px); the fp32 -> bf16 cast is fused here.
 * row_idx (device int64 [B], optional): output sample b reads image row_idx[b] — the minibatch gather of
 * PPOAgent.update (a10) is fused into this pass, so no gathered copy of the observations is ever made. */
int sc_patchify_bf16(const void* img, int img_is_bf16, void* out, int64_t B, int C, int H, int W,
                     int P, const int64_t* row_idx, void* stream);

/* LayerNorm over the last dim D (D % 256 == 0, D <= 1024).  io_f32: 0 = bf16 x/y, 1 = fp32 x/y.
 * mean/rstd: fp32 [M] saved for backward (may be null in fwd for inference). */
int sc_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                     float* rstd, int64_t M, int D, float eps, int io_f32, void* stream);
/* dx = dres + LN'(dy); dgamma/dbeta (+=, fp32 [D]); colsum (+=, fp32 [D], optional) gets the
 * column sums of the OUTPUT dx (the bias gradient of the linear layer that produced x). */
int sc_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
                     const float* gamma, const void* dres, void* dx, float* dgamma, float* dbeta,
                     float* colsum, int64_t M, int D, int io_f32, void* stream);

/* out[n] += sum_m x[m*ld + n]   (bias gradients); x bf16 */
int sc_colsum_bf16(const void* x, float* out, int64_t M, int N, int64_t ld, void* stream);

/* a3+a4: power normalisation + channel, fused, in-kernel Philox4x32-10 + Box-Muller.
 * z (fp32) is addressed as z[b*z_sb + m*z_sm + j], y likewise with (y_sb, y_sm); b<B, m<2, j<k.
 * scale: fp32 [B,2] = rsqrt(mean z^2 + eps), saved for backward.
 * snr_db: device fp32 [B] or null (then snr_db_scalar is used).
 * noise / fading: optional explicit tensors (fp32 [B,2,k] / [B,2,2]) replacing the Philox draw.
 * y_f32 / y_bf16: either may be null.  Counter layout: see oracle/philox.py.
 * rician_k: linear K-factor of SC_CH_RICIAN (>= 0; ignored by the other kinds; K = 0 is Rayleigh).  An explicit
 * `fading` tensor is the channel coefficient h itself for both fading kinds.
 * sample_idx (device int64 [B], optional): GLOBAL sample index of local sample b for the Philox counters (default
 * sample_offset + b) — a data-parallel rank that holds an arbitrary subset of the global minibatch draws the same
 * noise as the 1-rank run.
 * quant_bits > 0: the normalised symbols pass a 2^bits-level mid-rise uniform quantiser on [-quant_clip, quant_clip]
 * before the channel (oracle/channel.py: quantize_ste); backward is the clipped straight-through
 * estimator. */
int sc_pnorm_channel_fwd(const float* z, int64_t z_sb, int64_t z_sm, float* y_f32, void* y_bf16,
                         int64_t y_sb, int64_t y_sm, float* scale, int64_t B, int k, int kind, float rician_k,
                         const float* snr_db, float snr_db_scalar, uint64_t seed,
                         int64_t sample_offset, const int64_t* sample_idx, const float* noise, const float* fading,
                         float eps, const uint64_t* seed_dev, int quant_bits, float quant_clip, void* stream);
/* dz = scale * (g - zn * mean(g * zn)),  g = dy (AWGN/none) or |h|^2/(|h|^2+sigma^2) dy (Rayleigh / Rician) */
int sc_pnorm_channel_bwd(const float* dy, int64_t y_sb, int64_t y_sm, const float* z, int64_t z_sb,
                         int64_t z_sm, const float* scale, float* dz_f32, void* dz_bf16, int64_t B,
                         int k, int kind, float rician_k, const float* snr_db, float snr_db_scalar, uint64_t seed,
                         int64_t sample_offset, const int64_t* sample_idx, const float* fading,
                         const uint64_t* seed_dev, int quant_bits, float quant_clip, void* stream);
/* raw Philox words for parity tests: out u32 [n_quads,4], counters quad_offset.. */
int sc_philox_words(uint32_t* out, int64_t n_quads, uint32_t stream_id, uint64_t seed,
                    int64_t quad_offset, void* stream);
/* the in-kernel normal generator itself (stream 0, Box-Muller on the words above): out fp32 [n_quads, 4];
 * exact != 0 selects the accurate logf / sincospif form, 0 the bounded-error SFU form the channel kernel uses */
int sc_philox_normals(float* out, int64_t n_quads, uint64_t seed, int64_t quad_offset, int exact, void* stream);

/* a5: mean over S tokens: x bf16 [B,S,D] -> out_bf16 [B,D] (and fp32 copy, optional) */
int sc_meanpool_fwd(const void* x, void* out_bf16, float* out_f32, int64_t B, int S, int D,
                    void* stream);

/* a6: semantic distortion, fused forward + backward:
 *   loss_accum[0] += sum((r - t)^2) * inv_numel ;  dr = bf16(2 (r - t) * grad_scale)   (dr may be null)
 * r: bf16, numel elements.  t1 == null: t has the same flat layout as r.  t1 != null: r is
 * [samples, 2*seg_rows, D] (decoder tokens, modality-major per sample) and the target of rows
 * [0,seg_rows) is t[sample*seg_rows + row], of rows [seg_rows, 2*seg_rows) is t1[...] — the two
 * encoders' feature matrices, so no concatenated copy is ever made. */
int sc_mse_fwd_bwd(const void* r, const void* t, const void* t1, int seg_rows, int D,
                   float* loss_accum, void* dr, int64_t numel, float inv_numel, float grad_scale,
                   void* stream);

/* a7: actor-critic output layer on the tanh trunk: head_out[b, 0..A] = trunk[b,:] W^T + bias
 * trunk bf16 [B,D]; W fp32 [A+1, D]; head_out fp32 [B, A+1] (cols 0..A-1 logits, col A value) */
int sc_heads_fwd(const void* trunk, const float* W, const float* bias, float* head_out, int64_t B,
                 int D, int A1, void* stream);
/* dtrunk_pre = (dho W) * (1 - trunk^2) (bf16) ; dW += dho^T trunk ; dbias += colsum(dho) */
int sc_heads_bwd(const float* dho, const void* trunk, const float* W, void* dtrunk_pre, float* dW,
                 float* dbias, int64_t B, int D, int A1, void* stream);
/* a7: inverse-CDF sampling from supplied uniforms (fp64 cdf): action int64, logp fp32 */
int sc_sample_action(const float* head_out, int ld, const float* u, int64_t* action, float* logp,
                     int64_t B, int A, void* stream);

/* a8: GAE reverse scan, warp-shuffle linear-recurrence scan over T.
 * rewards/values/dones fp32 [T,N], last_value [N] -> adv, ret [T,N]; sums (double[2]) += (sum, sumsq) of adv */
int sc_gae(const float* rewards, const float* values, const float* dones, const float* last_value,
           float* adv, float* ret, double* sums, int T, int64_t N, float gamma, float lam,
           void* stream);
/* adv = (adv - mean) / (std + eps) with mean/std from sums over `count` elements */
int sc_adv_normalize(float* adv, const double* sums, int64_t count, int64_t n, float eps,
                     void* stream);

/* a9: PPO clipped loss, forward + backward fused.
 * head_out fp32 [M, ld] (A logits + value); writes dho fp32 [M, ld] (d total / d head_out, incl.
 * the 1/M_global factor) and accumulates stats[8] += {pi, v, ent, kl, clipfrac, 0,0,0} * inv_m
 * idx (device int64 [M], optional): sample i reads actions / logp_old / adv / ret at idx[i] (the minibatch gather of
 * PPOAgent.update fused into the loss; head_out and dho stay indexed by i). */
int sc_ppo_loss(const float* head_out, int ld, const int64_t* actions, const float* logp_old,
                const float* adv, const float* ret, const int64_t* idx, float* dho, float* stats, int64_t M, int A,
                float clip_eps, float c_v, float c_ent, float inv_m, void* stream);

/* a10: fused Adam over the flat parameter bucket; also refreshes the bf16 shadow copy.
 * g is multiplied by grad_scale (1/world_size) first. step >= 1. */
/* step_dev (optional, device int64): the effective step is step + *step_dev, read in the kernel, so a captured CUDA
 * graph advances the bias correction on every replay.  seed_dev of the channel kernels works the same way: the
 * Philox key is seed + *seed_dev. */
int sc_adam_flat(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr,
                 float beta1, float beta2, float eps, int64_t step, float grad_scale, const int64_t* step_dev,
                 void* stream);

int sc_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream);

/* a10 (PPOAgent.update): dst[i, :] = src[idx[i], :], rows of row_bytes (multiple of 16; idx NULL = plain copy).  src may
 * be PINNED HOST memory (unified addressing): the kernel is then the host -> device transfer of one minibatch — only its
 * rows cross PCIe, in permutation order, on a side stream while the previous optimizer step computes.  Grid: up to one
 * CTA of 512 threads per SM, max_ctas > 0 caps it (16 saturate PCIe and displace the fewest CTAs of a step running
 * beside it).  Indices outside [0, src_rows) leave their destination row untouched. */
int sc_gather_rows(const void* src, int64_t src_rows, int64_t row_bytes, const int64_t* idx, void* dst, int64_t n_rows,
                   int max_ctas, void* stream);

/* a10: first kernel of an optimizer step: zero the flat gradient bucket (n floats, n % 4 == 0) and the statistics
 * vector (n_stats floats), and advance the device-side step counter (*step_dev += 1, optional) that the channel seed
 * and Adam's bias correction read inside a captured CUDA graph — one launch instead of two memsets and an add. */
int sc_step_begin(float* grad, int64_t n, float* stats, int n_stats, int64_t* step_dev, void* stream);

/* SURVEY.md 8f rank 3, SNR-conditioned encoders (oracle/model.py: SnrEmbed): per-sample embedding of the channel SNR
 *   h[b, e] = tanh(s_b u[e] + v[e]),  emb[b, d] = sum_e W[d, e] h[b, e],  s_b = snr_db[b] / 20 (or the scalar)
 * added to every token of sample b by the patch-embedding GEMM (rowgrp epilogue).  u, v fp32 [E], W fp32 [D, E],
 * emb fp32 [B, D], h fp32 [B, E] (saved for backward).  E <= 256, D <= 1024.
 * bwd: dW += demb^T h, du += sum_b dh s_b, dv += sum_b dh with dh = (demb W) (1 - h^2); demb is multiplied by
 * demb_scale first (the caller passes the per-sample MEAN of the token gradients and demb_scale = tokens). */
int sc_snr_embed_fwd(const float* snr_db, float snr_db_scalar, const float* u, const float* v, const float* W,
                     float* emb, float* h, int64_t B, int D, int E, void* stream);
int sc_snr_embed_bwd(const float* demb, const float* h, const float* snr_db, float snr_db_scalar, const float* W,
                     float* dW, float* du, float* dv, float demb_scale, int64_t B, int D, int E, void* stream);

/* ---------------------------------------------------------------------------------------
 * SURVEY.md 8e / 8f rank 2 (ABI version 4): the data-parallel gradient exchange FUSED with the optimizer, over
 * NVLink peer memory instead of an NCCL all-reduce followed by sc_adam_flat.
 *
 * Every rank (one process per GPU) keeps its flat fp32 gradient bucket, fp32 master parameters and bf16 shadow in
 * PEER memory: allocated by sc_peer_alloc (plain cudaMalloc, not a pool, so it can be exported), exported as a
 * 64-byte handle (cudaIpcMemHandle_t), exchanged by the host (torch.distributed.all_gather_object) and mapped by
 * every other rank with sc_peer_import.  One optimizer step is then
 *
 *     sc_dp_barrier      every rank's backward is complete (its gradients are final, nobody reads parameters)
 *     sc_dp_reduce_adam  rank r owns slice r of the flat bucket (contiguous, 4-float granularity): it READS that slice
 *                        of every rank's gradients through NVLink, sums them in rank order 0..world-1 (every rank holds
 *                        bit-identical parameters afterwards), applies Adam with ITS moments (m / v are only
 *                        maintained for the own slice: optimizer state is sharded) and WRITES the updated fp32
 *                        parameters and the bf16 shadow of the slice into every rank's buffers
 *     sc_dp_barrier      every rank's writes have landed (parameters complete, gradient buckets may be zeroed)
 *
 * = reduce-scatter + sharded Adam + all-gather in one pass over the bucket; per rank (world - 1) / world of the
 * bucket crosses NVLink as loads and 1.5 x that as stores.  All three launches are CUDA-graph capturable.
 * These five sc_peer_* helpers are the only entry points that allocate or map memory.
 *
 * flags: peer memory, SC_DP_MAX_WORLD uint64 per rank, zero-initialised; flags[p] (host array of `world` DEVICE
 * pointers, entry `rank` being the local buffer) is rank p's flag array as mapped into this process.  epoch: local
 * device int64, zero-initialised, advanced by every barrier (all ranks call the same sequence of barriers).
 * A barrier that is not completed within ~20 s traps (the launch fails instead of hanging the GPU). */
#define SC_DP_MAX_WORLD 16
#define SC_PEER_HANDLE_BYTES 64
int sc_peer_alloc(void** ptr, int64_t bytes);                 /* zero-filled; device-synchronous */
int sc_peer_free(void* ptr);
int sc_peer_export(const void* ptr, void* handle);            /* handle: SC_PEER_HANDLE_BYTES host bytes */
int sc_peer_import(const void* handle, void** ptr);           /* maps another PROCESS's allocation */
int sc_peer_close(void* ptr);
int sc_dp_barrier(void* const* flags, int64_t* epoch, int world, int rank, void* stream);
/* g32 / p32 / p16: host arrays of `world` device pointers (entry `rank` = the local buffers); m, v local, n floats
 * each (n % 4 == 0).  step / step_dev as in sc_adam_flat.  max_ctas > 0 caps the grid (tests that run two ranks on
 * one device). */
int sc_dp_reduce_adam(const void* const* g32, void* const* p32, void* const* p16, float* m, float* v, int world,
                      int rank, int64_t n, float lr, float beta1, float beta2, float eps, int64_t step,
                      const int64_t* step_dev, int max_ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEMCOM_SM100_H_ */
