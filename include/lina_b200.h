/*
 * lina_b200.h -- C ABI of liblina_b200.so, the B200 (sm_100a) implementation of
 * Lina-Speech's GLA hot path and WavTokenizer decode tail.
 *
 * Every entry point replaces one function of the reference's operator API
 * (paths relative to the reference checkout; FLA/ = 3rdparty/flash-linear-attention/,
 * DEC/ = 3rdparty/decoder/).  The reference binds these from Python; the ctypes
 * stub a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions (all entry points):
 *   - plain C: pointers are DEVICE pointers, row-major contiguous, sizes are ints;
 *   - `dtype` is the element type of the activation tensors: LINA_F32 / LINA_BF16 / LINA_F16;
 *     all arithmetic is fp32 inside the kernels, final recurrent states are always fp32;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises,
 *     nothing allocates -- scratch memory is passed by the caller, sized by lina_*_workspace_bytes;
 *   - return value 0 = ok, negative = error (lina_last_error_string() describes the last
 *     error of the calling thread); no exceptions cross the boundary;
 *   - re-entrant and thread-safe (no global mutable state besides the per-thread error string).
 */
#ifndef LINA_B200_H
#define LINA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { LINA_F32 = 0, LINA_BF16 = 1, LINA_F16 = 2 };
enum {
    LINA_OK = 0,
    LINA_ERR_BAD_ARG = -1,      /* null pointer / non-positive size / unknown dtype */
    LINA_ERR_UNSUPPORTED = -2,  /* shape outside the implemented envelope (message says which) */
    LINA_ERR_CUDA = -3          /* a CUDA runtime call failed (message carries cudaGetErrorString) */
};

int         lina_abi_version(void);
const char *lina_last_error_string(void);

/* ---------------------------------------------------------------------------------------------
 * GLA recurrence.   S_t = diag(exp(gk_t)) S_{t-1} + k_t^T v_t ;   o_t = scale * q_t S_t
 *   q, k, gk : [B,H,T,K]   v, o : [B,H,T,V]   h0, ht : [B,H,K,V]
 * Replaces fla.ops.gla.fused_recurrent_gla  (FLA/fla/ops/gla/recurrent_fuse.py:13-27 ->
 * FLA/fla/ops/common/fused_recurrent.py:261-303) and is the spec of
 * fla.ops.gla.naive.naive_recurrent_gla (FLA/fla/ops/gla/naive.py:13-44).
 *   h0 may be NULL (zeros); its element type is h0_dtype.  ht may be NULL (not wanted), else fp32.
 * ------------------------------------------------------------------------------------------- */
int lina_gla_recurrent_fwd(const void *q, const void *k, const void *v, const void *gk,
                           const void *h0, int h0_dtype, void *o, float *ht,
                           int B, int H, int T, int K, int V, int dtype, float scale, void *stream);

/* Backward of the above (FLA/fla/ops/common/fused_recurrent.py:305-343, kernel :118-257):
 * dq, dk, dgk [B,H,T,K] and dv [B,H,T,V] in `dtype`; dh0 [B,H,K,V] fp32 (NULL = not wanted);
 * dht [B,H,K,V] fp32 or NULL.  `ws` is scratch of lina_gla_recurrent_bwd_workspace_bytes(). */
size_t lina_gla_recurrent_bwd_workspace_bytes(int B, int H, int T, int K, int V);
int lina_gla_recurrent_bwd(const void *q, const void *k, const void *v, const void *gk,
                           const void *h0, int h0_dtype, const void *d_o, const float *dht,
                           void *dq, void *dk, void *dv, void *dgk, float *dh0, void *ws,
                           int B, int H, int T, int K, int V, int dtype, float scale, void *stream);

/* RWKV6 variant of the recurrence, forward only (secondary row a13: reachable only from the reference's stale
 * model/rwkv6.py).  o_t = scale * r_t (S_{t-1} + diag(u) k_t^T v_t) ; S_t = diag(exp(w_t)) S_{t-1} + k_t^T v_t, u [H,K].
 * Replaces fla.ops.rwkv6.fused_recurrent_rwkv6 / chunk_rwkv6 forward (FLA/fla/ops/rwkv6/recurrent_fuse.py:335-368,
 * FLA/fla/ops/rwkv6/chunk.py:803-); spec FLA/fla/ops/rwkv6/recurrent_naive.py:8-42. */
int lina_rwkv6_recurrent_fwd(const void *r, const void *k, const void *v, const void *w, const void *u,
                             const void *h0, int h0_dtype, void *o, float *ht,
                             int B, int H, int T, int K, int V, int dtype, float scale, void *stream);

/* Chunkwise-parallel forward of the same function -- the tensor-core path.
 * Replaces fla.ops.gla.fused_chunk_gla (FLA/fla/ops/gla/chunk_fuse.py:518-536) and
 * fla.ops.gla.chunk_gla (FLA/fla/ops/gla/chunk.py:453-491); same contract as
 * lina_gla_recurrent_fwd, arbitrary T.  `ws` is scratch of lina_gla_chunk_fwd_workspace_bytes(). */
size_t lina_gla_chunk_fwd_workspace_bytes(int B, int H, int T, int K, int V, int dtype);
int lina_gla_chunk_fwd(const void *q, const void *k, const void *v, const void *gk,
                       const void *h0, int h0_dtype, void *o, float *ht, void *ws,
                       int B, int H, int T, int K, int V, int dtype, float scale, void *stream);
/* Same function on the layout the projections produce: q, k, gk [B,T,H,K], v, o [B,T,H,V] (h0 / ht stay
 * [B,H,K,V]).  Replaces the `rearrange(..., 'b l (h d) -> b h l d')` + `.contiguous()` copies in front of the
 * op (model/gla.py:173, FLA/fla/utils.py:9-15) and the inverse rearrange after it (model/gla.py:215): TMA reads
 * the strided view in place.  Tensor-core envelope only (see lina_gla_chunk_fwd_uses_tensor_cores), else
 * LINA_ERR_UNSUPPORTED. */
int lina_gla_chunk_fwd_bthd(const void *q, const void *k, const void *v, const void *gk,
                            const void *h0, int h0_dtype, void *o, float *ht, void *ws,
                            int B, int H, int T, int K, int V, int dtype, float scale, void *stream);
/* 1 if lina_gla_chunk_fwd runs the tcgen05 (sm_100a tensor-core) kernel for this problem,
 * 0 if it runs the CUDA-core recurrence kernel (small/odd shapes, fp32 inputs). */
int lina_gla_chunk_fwd_uses_tensor_cores(int B, int H, int T, int K, int V, int dtype);

/* ---------------------------------------------------------------------------------------------
 * One autoregressive step (T = 1) of a whole GLA mixer between its GEMMs, state updated IN PLACE.
 * Replaces, for GatedLinearAttention.forward with a cache and one token (model/gla.py:146-220):
 *   3x ShortConvolution.step (FLA/fla/modules/convolution.py:180-205), logsigmoid/normaliser
 *   (model/gla.py:174-176), the GLA op at T = 1 (model/gla.py:187-193), Cache.update's copy_
 *   (FLA/fla/models/utils.py:61-66) and FusedRMSNormSwishGate (FLA/fla/modules/fused_norm_gate.py:72-139).
 *   xq, xk : [B, H*K]   xv, g : [B, H*V]     projections of the token (dtype)
 *   gk_raw : [B, H*K]   gate logits before logsigmoid (dtype)
 *   wq, wk : [H*K, W]   wv : [H*V, W]        depthwise taps (dtype); NULL = no short conv
 *   cq, ck : [B, H*K, W]  cv : [B, H*V, W]   conv states (state_dtype), rolled in place
 *   S      : [B, H, K, V]                    recurrent state (state_dtype), updated in place
 *   norm_w : [V] (dtype)                     out : [B, H*V] (dtype) = RMSNorm(o) * w * swish(g)
 *   ws     : scratch of lina_gla_step_workspace_bytes()
 * ------------------------------------------------------------------------------------------- */
size_t lina_gla_step_workspace_bytes(int B, int H, int K, int V);
int lina_gla_step(const void *xq, const void *xk, const void *xv, const void *gk_raw, const void *g,
                  const void *wq, const void *wk, const void *wv, void *cq, void *ck, void *cv,
                  void *S, const void *norm_w, void *out, void *ws,
                  int B, int H, int K, int V, int W, int dtype, int state_dtype,
                  float scale, float gate_normalizer, float eps, void *stream);

/* Same, reading xq, xk, xv, g as column slices of ONE projection buffer with row stride `ldx` elements (the
 * output of a single [q;k;v;g;..] GEMM) and gk_raw with row stride `ldg`; 0 = dense.  Saves the split copies. */
int lina_gla_step_ld(const void *xq, const void *xk, const void *xv, const void *gk_raw, const void *g,
                     const void *wq, const void *wk, const void *wv, void *cq, void *ck, void *cv,
                     void *S, const void *norm_w, void *out, void *ws,
                     int B, int H, int K, int V, int W, int dtype, int state_dtype,
                     float scale, float gate_normalizer, float eps, int ldx, int ldg, void *stream);

/* Same, with the gate logits computed inside from the rank-R factors of gk_proj (model/gla.py:96-97): lo [B,R] (row stride
 * ld_lo), w2 [H*K,R], b2 [H*K] or NULL -- one GEMM launch less per layer per token. */
int lina_gla_step_lr(const void *xq, const void *xk, const void *xv, const void *lo, int ld_lo, const void *w2,
                     const void *b2, int R, const void *g, const void *wq, const void *wk, const void *wv,
                     void *cq, void *ck, void *cv, void *S, const void *norm_w, void *out, void *ws,
                     int B, int H, int K, int V, int W, int dtype, int state_dtype,
                     float scale, float gate_normalizer, float eps, int ldx, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Post-projection pass of a whole-sequence GLA mixer call in ONE launch (inference prefill):
 *   q, k, v = SiLU(ShortConvolution(x_q | x_k | x_v))   (FLA/fla/modules/convolution.py:141-178, model/gla.py:161-163)
 *   gk      = logsigmoid(gk_raw) / normalizer [clamped]   (model/gla.py:174-181)
 *   xq, xk [B,L,Dk], xv [B,L,Dv] with row stride `ldx` elements (column slices of one [q;k;v;g] GEMM output);
 *   gk_raw [B,L,Dk] with row stride `ldg`; wq, wk [Dk,4], wv [Dv,4]; outputs contiguous [B,L,D];
 *   cq, ck [B,Dk,4], cv [B,Dv,4] receive the last 4 inputs (all three or none).  W must be 4.
 * ------------------------------------------------------------------------------------------- */
int lina_gla_prefill_prep(const void *xq, const void *xk, const void *xv, long long ldx,
                          const void *wq, const void *wk, const void *wv, const void *gk_raw, long long ldg,
                          void *q, void *k, void *v, void *gk, void *cq, void *ck, void *cv, int cache_dtype,
                          int B, int L, int Dk, int Dv, int W, float gate_normalizer, float clamp_min, int use_clamp,
                          int dtype, void *stream);

/* Rank-R expansion linear out[m, n] = bias[n] + sum_{r<R} x[m, r] W[n, r] (bf16, fp32 accumulate; R in {8, 16, 32}; bias nullable;
 * x rows `ldx` elements apart, out rows `ldo`): gk_proj[1] of GatedLinearAttention (model/gla.py:96-97,
 * nn.Linear(gate_low_rank_dim, key_dim)) over a whole sequence -- an HBM-write-bound outer-product expansion that a library GEMM
 * (K = 16) serves 10x off its bandwidth bound. */
int lina_lowrank_linear(const void *x, long long ldx, const void *W, const void *bias, void *out, long long ldo,
                        int M, int N, int R, int dtype, void *stream);

/* Same pass with the chunk gating of the tensor-core GLA kernel folded in (bf16 only): instead of q, k, gk it writes
 *   qg = scale * q * e^G,  kg = k * e^-G   [B,L,H*K] bf16     (G = cumsum of gk inside each 64-token chunk, fp32)
 *   decay[b,h,n,:] = e^{G at the chunk end}  [B,H,ceil(L/64),K] fp32
 * i.e. the MMA operands FLA/fla/ops/gla/chunk_util.py:28-65 (prepare_qg_kg) materialises, for lina_gla_chunk_fwd_pregated_bthd.
 * v = SiLU(ShortConvolution(x_v)) as above; xq / xk / xv have their own row strides (separate or concatenated GEMMs).
 * The gate normalizer must be a power of two (16 in the shipped model).
 * ``envelope_flag`` (device int, nullable): OR-ed with 1 when the summed log gate of any (chunk, channel) falls below -80,
 * i.e. outside the range the single-pivot tensor-core kernel represents (the reference is exact for any gate,
 * FLA/fla/ops/gla/chunk_fuse.py:238-247 computes the intra-chunk scores with per-pair exponent differences): the caller
 * must then serve the call through lina_gla_prefill_prep + the exact recurrence instead. */
int lina_gla_prefill_prep_gated(const void *xq, long long ldq, const void *xk, long long ldk, const void *xv, long long ldv,
                                const void *wq, const void *wk, const void *wv, const void *gk_raw, long long ldg,
                                void *qg, void *kg, void *v, float *decay, void *cq, void *ck, void *cv, int cache_dtype,
                                int B, int L, int H, int K, int V, int W, float gate_normalizer, float scale,
                                int *envelope_flag, void *stream);
/* The chunkwise GLA forward (lina_gla_chunk_fwd_bthd) on those pre-gated operands: no in-kernel gate pre-pass, gk is
 * never read.  Tensor-core envelope only. */
int lina_gla_chunk_fwd_pregated_bthd(const void *qg, const void *kg, const void *v, const float *decay,
                                     const void *h0, int h0_dtype, void *o, float *ht,
                                     int B, int H, int T, int K, int V, void *stream);
/* Same with a device workspace (256-byte aligned, lina_gla_chunk_fwd_pregated_ws_bytes(); 0 = none needed).  The kernel runs
 * (batch, head, pair of 128-wide V slices) tiles on CTA pairs; when the last wave of tiles would leave more than half of the
 * GPU idle (512 V slices on 148 SMs: 3.46 waves), those tiles are cut in two along T -- the first halves run first, leave
 * their fp32 state slice in the workspace, the second halves run last -- the role FLA/fla/ops/common/chunk_h.py:15-98 gives NT
 * in its grid.  Results are bit-identical to the call without workspace. */
size_t lina_gla_chunk_fwd_pregated_ws_bytes(int B, int H, int T, int K, int V);
int lina_gla_chunk_fwd_pregated_bthd_ws(const void *qg, const void *kg, const void *v, const float *decay,
                                        const void *h0, int h0_dtype, void *o, float *ht, void *ws, size_t ws_bytes,
                                        int B, int H, int T, int K, int V, void *stream);

/* General form of the same kernel: [B,H,T,D] (bthd = 0) or [B,T,H,D] (bthd = 1) operands; row_decay != 0: `decay` is
 * [B,H,NT,V] and scales the VALUE dim of the state, out_f32 != 0: o is fp32 (only (0,0) and (1,1) are instantiated).
 * The chunked BACKWARD of the GLA op (replacing FLA/fla/ops/gla/chunk.py:140-341 + FLA/fla/ops/common/chunk_h.py:111-189)
 * is five runs of this kernel on role-swapped, time-reversed operands -- see lina_speech_b200/fla_api/ops.py:_bwd_tc. */
int lina_gla_chunk_fwd_pregated(const void *qg, const void *kg, const void *v, const float *decay,
                                const void *h0, int h0_dtype, void *o, float *ht,
                                int B, int H, int T, int K, int V, int bthd, int row_decay, int out_f32,
                                int ldq, int ldk, int ldv, void *stream);
/* Element-wise passes of that backward (bf16; [B,H,T,K] contiguous or, with bthd != 0, [B,T,H,K]; chunk 64; Tp = T rounded up to whole chunks):
 *   prep  : kt [B,H,T,K] = k e^-G ; qh_r, kh_r [B,H,Tp,K] = scale q e^{G-G_C}, k e^{G_C-G} time-reversed (zero rows first when
 *           T is ragged) ; D, Dr [B,H,NT,K] fp32 = e^{G_C} in forward / reversed chunk order.
 *   time_reverse_pad2 : a_r[bh, Tp-1-t] = a[bh, t] (zero rows for t >= T), same for b; rows of Dm elements.
 *   post  : dq = (dqa + dqb) scale e^G, dk = (dka + dkb)[reversed] e^{G_C-G} (dqb / dkb may be NULL), dgk_local = in-chunk
 *           reversed cumsum of dq q - dk k (fp32), totals [B,H,NT,K] its chunk sums.
 *   dgk_finish : dgk = bf16(dgk_local + carry[b,h,t/64,:]). */
int lina_gla_bwd_prep(const void *q, const void *k, const void *gk, void *kt, void *qh_r, void *kh_r, float *D, float *Dr,
                      int B, int H, int T, int K, int bthd, float scale, void *stream);
int lina_time_reverse_pad2(const void *a, const void *b, void *a_r, void *b_r, long long BH, int T, int Tp, int Dm,
                           void *stream);
int lina_gla_bwd_post(const float *dqa, const float *dqb, const float *dka, const float *dkb, const void *q, const void *k,
                      const void *gk, void *dq, void *dk, float *dgk_local, float *totals,
                      int B, int H, int T, int K, int bthd, float scale, void *stream);
int lina_gla_bwd_dgk_finish(const float *dgk_local, const float *carry, void *dgk, int B, int H, int T, int K, int bthd,
                            void *stream);

/* ---------------------------------------------------------------------------------------------
 * ShortConvolution: y[b,l,d] = act(sum_j w[d,j] * x[b, l-(W-1)+j, d]), act = SiLU or identity.
 * Replaces causal_conv1d_fn / causal_conv1d_update (causal-conv1d 1.3.0.post1, call sites
 * FLA/fla/modules/convolution.py:168-173,189-195; semantics = the torch branch :175-178,197-204).
 *   x, y : [B,L,D]   w : [D,W]   cache : [B,D,W] or NULL (receives the last W inputs, :164-166).
 * ------------------------------------------------------------------------------------------- */
int lina_short_conv_fwd(const void *x, const void *w, void *y, void *cache, int cache_dtype,
                        int B, int L, int D, int W, int silu, int dtype, void *stream);
int lina_short_conv_bwd(const void *x, const void *w, const void *dy, void *dx, float *dw,
                        int B, int L, int D, int W, int silu, int dtype, void *stream);
int lina_short_conv_update(const void *x, void *cache, int cache_dtype, const void *w, void *y,
                           int B, int D, int W, int silu, int dtype, void *stream);

/* ---------------------------------------------------------------------------------------------
 * FusedRMSNormSwishGate: y = x * rsqrt(mean(x^2) + eps) * w * g * sigmoid(g), rows of length N.
 * Replaces rms_norm_swish_gate_fn (FLA/fla/modules/fused_norm_gate.py:439-518; kernels :72-139, :220-335).
 *   x, g, y : [M,N]   w : [N] or NULL   rstd : [M] fp32 (saved for backward; may be NULL in fwd).
 *   bwd: dw is an fp32 [N] accumulator that must be zeroed by the caller.
 * ------------------------------------------------------------------------------------------- */
int lina_rmsnorm_swishgate_fwd(const void *x, const void *g, const void *w, void *y, float *rstd,
                               int M, int N, float eps, int dtype, void *stream);
/* Same forward with the gate read in place from a wider buffer: row r of g starts at
 * g + (r / g_group) * ldg + (r % g_group) * N  (g_group = heads, ldg = row stride of the [q;k;v;g] projection). */
int lina_rmsnorm_swishgate_fwd_ld(const void *x, const void *g, const void *w, void *y, float *rstd,
                                  int M, int N, float eps, int g_group, long long ldg, int dtype, void *stream);
int lina_rmsnorm_swishgate_bwd(const void *x, const void *g, const void *w, const float *rstd,
                               const void *dy, void *dx, void *dg, float *dw,
                               int M, int N, int dtype, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Fused element-wise passes of the block around the mixer.
 *   gate : y = logsigmoid(x) / normalizer, optionally clamped from below     (model/gla.py:174-181), n elements
 *   swiglu: out[m, j] = silu(h[m, j]) * h[m, Hp + j], h [M, 2*Hp] -> out [M, Hp]  (model/base_blocks.py:48-50)
 * ------------------------------------------------------------------------------------------- */
int lina_gate_logsigmoid(const void *x, void *y, long long n, float normalizer, float clamp_min, int use_clamp,
                         int dtype, void *stream);
int lina_swiglu_act(const void *h, void *out, int M, int Hp, int dtype, void *stream);
/* sum_out = a + x ; ln_out = LayerNorm(sum_out) * gamma + beta over rows of length N (nn.LayerNorm semantics).
 * a == NULL: plain LayerNorm of x (sum_out unused).  The residual-add + pre-LN pairs of MixingBlock
 * (model/base_blocks.py:65-68) in one pass. */
int lina_add_layernorm(const void *a, const void *x, const void *gamma, const void *beta, void *sum_out,
                       void *ln_out, int M, int N, float eps, int dtype, void *stream);

/* nn.LayerNorm for the bf16-autocast TRAINING path (MixingBlock norm1 / norm2, model/base_blocks.py:65-68): fp32 residual
 * stream in, `out_dtype` output (what every consuming autocast Linear would cast it to anyway), mean / rstd [M] saved.
 * Backward: dx fp32, dgamma / dbeta fp32 accumulators that the caller zeroes.  N % 4 == 0, N <= 1024. */
int lina_layernorm_f32in_fwd(const float *x, const float *gamma, const float *beta, void *y, float *mean, float *rstd,
                             int M, int N, float eps, int out_dtype, void *stream);
int lina_layernorm_f32in_bwd(const float *x, const float *gamma, const float *mean, const float *rstd, const void *dy,
                             float *dx, float *dgamma, float *dbeta, int M, int N, int dy_dtype, void *stream);
/* Row-wise cross entropy of LinaModel.forward (model/modeling_lina.py:104-106: F.cross_entropy(logits.float(),
 * target, ignore_index=1)) straight from the `dtype` logits, fp32 math: loss[m] = logsumexp(logits[m,:Vn]) -
 * logits[m,target[m]], valid[m] = 1; rows with target == ignore_index or row_mask[m] == 0 give loss = valid = 0.
 * logits rows are `ld` elements apart (>= Vn: a vocabulary-padded GEMM output is read in place); row_mask may be NULL.
 * The caller reduces: loss.sum() / valid.sum(). */
int lina_cross_entropy_rows(const void *logits, long long ld, const int64_t *target, const uint8_t *row_mask,
                            float *loss, float *valid, int M, int Vn, long long ignore_index, int dtype, void *stream);

/* Fused top-k sampling of the decode loop (model/tools.py:38-44 `topk_sampling`, called per quantizer at
 * model/modeling_lina.py:159-165): kth = k-th largest UNSCALED logit of the row, keep x/temp >= kth (the reference's quirk),
 * softmax over the kept entries, id = inverse-CDF sample with the caller's uniform[b] in [0,1).  logits [B, ld] (`dtype`,
 * first Vn columns), out [B] int64.  k = 1 returns the arg-max.  torch.multinomial's bit pattern is not reproduced (same
 * distribution). */
int lina_topk_sample(const void *logits, long long ld, int B, int Vn, int k, float temp, const float *uniform,
                     int64_t *out, int dtype, void *stream);

/* ---------------------------------------------------------------------------------------------
 * WavTokenizer decode tail (fp32).  The dense convolutions / linears of the backbone stay library
 * GEMMs on the host side; these are the HBM-bound stages between them.
 * ------------------------------------------------------------------------------------------- */
/* codes_to_features (DEC/pretrained.py:209-239): codes [K,B,L] int64, codebooks [K*bins, C] ->
 * features [B,C,L] (sum over the K codebooks, transposed). */
int lina_codec_codes_to_features(const int64_t *codes, const float *codebooks, float *features,
                                 int Kq, int B, int L, int bins, int C, void *stream);
/* GroupNorm(groups, eps, affine) + optional swish on [B,C,L] (DEC/models.py:10-16,58-70,113).
 * `ws` = 2*B*groups floats. */
int lina_codec_groupnorm_swish(const float *x, const float *gamma, const float *beta, float *y, float *ws,
                               int B, int C, int L, int groups, float eps, int swish, void *stream);
/* ConvNeXt front half (DEC/modules.py:45-50): depthwise Conv1d(k=7,pad=3) on [B,C,L] + bias, transpose
 * to [B,L,C], LayerNorm(no affine, eps) * scale[C] + shift[C] (AdaLayerNorm, DEC/modules.py:81-86).
 * dw_w NULL = skip the conv (plain transposing AdaLN, DEC/models.py:229). */
int lina_codec_dwconv_adaln(const float *x, const float *dw_w, const float *dw_b, const float *scale,
                            const float *shift, float *y, int B, int C, int L, float eps, void *stream);
/* Same two functions as a pair of small kernels (tile = 64 channels x 64 steps, deterministic merged statistics) instead of
 * one kernel holding all C channels of a time tile: `ws` = lina_codec_dwconv_adaln_workspace_bytes(B, C, L). */
size_t lina_codec_dwconv_adaln_workspace_bytes(int B, int C, int L);
int lina_codec_dwconv_adaln_ws(const float *x, const float *dw_w, const float *dw_b, const float *scale,
                               const float *shift, float *y, float *ws, int B, int C, int L, float eps, void *stream);
int lina_codec_layernorm_t_ws(const float *x, const float *gamma, const float *beta, float *y, float *ws,
                              int B, int C, int L, float eps, void *stream);
/* ConvNeXt back half (DEC/modules.py:55-59): out[b,c,l] = res[b,c,l] + gamma[c] * h[b,l,c]. */
int lina_codec_scale_residual_t(const float *h, const float *gamma, const float *res, float *out,
                                int B, int C, int L, void *stream);
/* LayerNorm over the channel dim of [B,C,L] written as [B,L,C] (final_layer_norm, DEC/models.py:234). */
int lina_codec_layernorm_t(const float *x, const float *gamma, const float *beta, float *y,
                           int B, int C, int L, float eps, void *stream);
/* ISTFTHead tail (DEC/heads.py:53-67 + DEC/spectral_ops.py:33-75, padding="same"):
 * h [B,L,n_fft+2] (Linear output: log-magnitudes then phases) -> wav [B, L*hop].
 * mag = min(exp(.),100); S = mag*(cos p + i sin p); irfft(n_fft) * window; overlap-add with hop;
 * trim (n_fft-hop)/2 per side; divide by the overlap-added window^2 envelope.
 * `ws` = lina_codec_istft_workspace_bytes(). */
size_t lina_codec_istft_workspace_bytes(int B, int L, int n_fft);
int lina_codec_istft_head(const float *h, const float *window, float *wav, void *ws,
                          int B, int L, int n_fft, int hop, void *stream);
/* Same with a row stride (elements) for h: the head GEMM writes rows padded to a multiple of 4 floats. */
int lina_codec_istft_head_ld(const float *h, long long ldh, const float *window, float *wav, void *ws,
                             int B, int L, int n_fft, int hop, void *stream);

/* Channels-last stages between the tensor-core contractions (csrc/codec_cl.cu); activations [B, L, C] fp32, outputs fp32
 * and / or `n_parts` bf16 parts (hi [, mid], lo) of the fp32 result -- the A operand of the next lina_gemm_bf16_terms.
 * codes -> sum over quantizers of codebook rows (DEC/pretrained.py:231-237), without the [B,C,L] feature tensor: */
int lina_codec_cl_gather(const int64_t *codes, const float *codebooks, float *out_f32, void *const *out_parts, int n_parts,
                         int Kq, int B, int L, int bins, int C, void *stream);
/* GroupNorm(G, C) statistics of [B, L, C] (DEC/models.py:15-16): Welford partials per (batch, 32-row tile, group),
 * merged by the consumer (lina_codec_cl_rows). */
size_t lina_codec_cl_gn_partials_bytes(int B, int L, int G);
int lina_codec_cl_gn_partials(const float *x, float *partials, int B, int L, int C, int G, void *stream);
/* One pass per row: [depthwise Conv1d k=7, padding 3 (dw_w [C,7], dw_b [C]; DEC/modules.py:28,48)] ->
 * [GroupNorm apply from the partials, affine gn_w / gn_b] -> [swish (DEC/models.py:10-12)] ->
 * [LayerNorm over C without affine, * ln_scale + ln_shift (AdaLayerNorm DEC/modules.py:81-86; nn.LayerNorm with its
 * weight / bias)].  Stages with a NULL first pointer are skipped. */
int lina_codec_cl_rows(const float *x, const float *dw_w, const float *dw_b, const float *gn_partials, const float *gn_w,
                       const float *gn_b, int G, float gn_eps, int swish, const float *ln_scale, const float *ln_shift,
                       float ln_eps, float *out_f32, void *const *out_parts, int n_parts, int B, int L, int C, void *stream);
/* softmax over the key axis of the attention scores S [rows, ldS] (row length n) -> bf16 parts [rows, ldP], columns
 * n..ldP-1 zero (DEC/models.py:119-120). */
int lina_codec_cl_softmax(const float *S, void *const *out_parts, int n_parts, long long rows, int n, long long ldS,
                          long long ldP, void *stream);

/* One decode step of a single-head softmax attention against memoised text-side tensors (BlindCrossAttention, model/crossatt.py:
 * 105-155, eval branch :13-19): w = softmax(LayerNorm?(q) K^T * scale), out = w V per sequence.  q [B, ldq]; keys / vals
 * [B, n, d] (batch stride in elements; 0 = one table shared by every sequence, e.g. the positional embeddings);
 * att_out [B, att_bstride] receives w (nullable); out [B, ldo].  ln_w / ln_b NULL = no LayerNorm on q. */
int lina_cross_att_step(const void *q, long long ldq, const void *ln_w, const void *ln_b, float eps, const void *keys,
                        long long key_bstride, const void *vals, long long val_bstride, void *att_out, long long att_bstride,
                        void *out, long long ldo, int B, int n, int d, float scale, int dtype, void *stream);

/* Linear layers of ONE autoregressive step (M = batch <= lina_skinny_linear_max_rows() rows, bf16): weight-streaming kernel
 * with the step's row-wise neighbours fused in (csrc/skinny_linear.cu).  Replaces, per MixingBlock and token
 * (model/base_blocks.py:65-69, model/gla.py:91-99,225), nn.LayerNorm + the residual add in front of a projection and
 * SwiGLU's silu(gate) * u behind it:
 *   s   = x + delta                      (only with delta; written to sum_out, [M, K] contiguous; delta [M, ldd])
 *   h   = LayerNorm(s) * ln_gamma + ln_beta   (only with ln_gamma / ln_beta; else h = x)
 *   y   = h W^T + bias                   W [N, ldw] row-major
 *   out = y                              (swiglu_pair_offset == 0)
 *   out[:, j] = silu(y[:, j]) * y[:, swiglu_pair_offset + j]   (> 0: W holds gate rows [0, N) and u rows [offset, offset + N))
 * Roundings follow the unfused sequence (bf16 after the add, after the LayerNorm, after the linear, after the product). */
int lina_skinny_linear_max_rows(void);
int lina_skinny_linear(const void *x, long long ldx, const void *delta, long long ldd, const void *ln_gamma, const void *ln_beta,
                       float ln_eps, void *sum_out, const void *W, long long ldw, const void *bias, void *out, long long ldo,
                       int M, int N, int K, int swiglu_pair_offset, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Tensor-core contractions of the WavTokenizer decoder at fp32 fidelity (csrc/gemm_sm100.cu): replaces the cuDNN / cuBLAS
 * fp32 calls behind nn.Conv1d k = 7 / k = 3 / 1x1 (DEC/models.py:177,203-216 -> 40-51,97-105), nn.Linear of ConvNeXtBlock
 * (DEC/modules.py:33-37,52-56) and ISTFTHead.out (DEC/heads.py:39,53), and the two torch.bmm of AttnBlock
 * (DEC/models.py:116-124).
 *
 *   D[b,l,n] = alpha * sum_tap sum_(pa,pb) sum_k A_pa[b, l + tap - pad, k] * B_pb[(b,) n, tap*K + k]
 *   out      = act(D + bias[n]) * gamma[n] + residual[b,l,n]
 *
 * Operands are bf16 "parts" of fp32 tensors (x = part0 + part1 [+ part2], part0 = bf16(x), part i = bf16 of the remainder);
 * `terms` lists the part products to accumulate (all in one fp32 accumulator): {(0,0)} is a plain bf16 GEMM,
 * {(0,0),(0,1),(1,0)} carries 16 significand bits per operand, the six-term list over three parts 24 (fp32).
 * A parts: [NB, L, lda] row-major (rows outside [0, L) of a batch read as zero: the convolution's zero padding);
 * B parts: [N, ldb] (weights, [N][tap][K]) or, with b_batched, [NB, N, ldb]; with b_mn the transposed storage [K, ldb]
 * (how the attention block's V sits in memory).  Strides in elements, multiples of 8.
 * Outputs: out_f32 [NB*L, ld_out] and / or out_parts bf16 parts [NB*L, ld_split] of the result (the next GEMM's A).
 * act: 0 none, 1 GELU (erf), 2 swish.  bias / gamma / residual may be NULL. */
typedef struct {
    const void *a[3];
    long long lda, a_batch_stride;      /* a_batch_stride 0 = L * lda */
    int a_parts;
    const void *b[3];
    long long ldb, b_batch_stride;      /* b_batch_stride 0 = N * ldb */
    int b_parts, b_batched;
    int n_terms, term_a[6], term_b[6];
    int NB, L, N, K, taps, pad;
    float alpha;
    const float *bias, *gamma, *residual;
    long long ld_res;
    int act;
    float *out_f32;
    long long ld_out;
    void *out_split[3];
    long long ld_split;
    int out_parts;
    int b_mn;                           /* 1: B parts stored [K, ldb] (N contiguous) / [NB, K, ldb]; taps must be 1 */
    int span;                           /* 64-wide K blocks accumulated on the tensor core between two promotions of the partial
                                           sums to fp32 registers (0 = default 8: the tensor core's accumulator truncates after
                                           every K = 16 step; a large span trades that error for fewer TMEM reads) */
} lina_gemm_args;
int lina_gemm_bf16_terms(const lina_gemm_args *args, void *stream);

/* A/B switches for kernel variants (bring-up only; process-global, not thread-safe): key 0 = rows per thread of the
 * prep / short-conv tile kernel (8 or 16), key 1 = 1 selects the round-1 sliding-window short-conv kernel,
 * key 2 = bit mask of tcgen05 GLA kernel options, key 3 = 1 selects the scalar-fp32 short-conv tile kernel for bf16,
 * key 4 = 1 runs the pre-gated GLA kernel's state pass on one warpgroup, key 5 = 1 selects the round-1 short-conv backward,
 * (keys 6 / 7 selected a three-stage operand ring and cluster-multicast operand loads of the pre-gated GLA kernel: both measured
 * slower or neutral in round 1 and no longer built), key 8 = 2 routes lina_codec_istft_head (n_fft = 1280) back to the generic shared-memory
 * FFT (default: the warp-per-frame fixed-radix kernel, csrc/fft640.cuh: 0.25 vs 0.69 ms at 32 x 750 frames), key 9 = 1 selects
 * the one-CTA-per-tile pre-gated GLA kernel of round 1 (2: CTA pairs without the T cut; default 0: pairs + T cut). */
int lina_debug_set_variant(int key, int value);

#ifdef __cplusplus
}
#endif
#endif /* LINA_B200_H */
