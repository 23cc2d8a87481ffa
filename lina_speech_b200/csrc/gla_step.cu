// One autoregressive step (T = 1) of a GLA mixer between its GEMMs, states updated in place.
//
// Replaces, per GatedLinearAttention.forward call with a cache and one token (model/gla.py:146-220):
// 3x causal_conv1d_update (FLA/fla/modules/convolution.py:180-205), logsigmoid/normaliser
// (model/gla.py:174-176), the T=1 GLA op (4 Triton launches + bmm + 2 sums in the default
// fused_chunk mode, FLA/fla/ops/gla/chunk_fuse.py:307-394), Cache.update's 4 copies
// (FLA/fla/models/utils.py:61-66) and the norm-gate kernel (FLA/fla/modules/fused_norm_gate.py:72-139).
//
// HBM-bound: the recurrent state S [B,H,K,V] is read once and written once (in place); the
// algorithmic bytes are 2 * B*H*K*V * sizeof(state) per call, everything else is < 2 % of that.
//   k1 gla_step_prep  : conv roll+dot+SiLU for q,k,v ; e = exp(logsigmoid(gk_raw)/normalizer)   (tiny)
//   k2 gla_step_state : S <- e (.) S + k^T v ; o = scale q S     128-bit loads/stores, 8 rows in flight/thread
//   k3 gla_step_norm  : RMSNorm(o) * w * swish(g)                                               (tiny)
#include "common.cuh"

namespace {

template <typename T, typename CT>
__global__ void gla_step_prep_kernel(const T *__restrict__ xq, const T *__restrict__ xk, const T *__restrict__ xv,
                                     const T *__restrict__ gk_raw, const T *__restrict__ wq,
                                     const T *__restrict__ wk, const T *__restrict__ wv, CT *__restrict__ cq,
                                     CT *__restrict__ ck, CT *__restrict__ cv, float *__restrict__ qf,
                                     float *__restrict__ kf, float *__restrict__ ef, float *__restrict__ vf,
                                     int B, int HK, int HV, int W, float scale, float inv_norm, int ld_qk, int ld_v, int ldg,
                                     const T *__restrict__ lo, int ld_lo, const T *__restrict__ w2,
                                     const T *__restrict__ b2, int R) {
    const int per_b = 3 * HK + HV;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * per_b) return;
    const int b = (int)(i / per_b);
    int c = (int)(i - (long long)b * per_b);
    if (c >= 2 * HK + HV) {                         // gate channel
        c -= 2 * HK + HV;
        float x;
        if (lo != nullptr) {
            // gate logits from the rank-R factorisation in place: x = b2[c] + sum_r lo[b, r] * w2[c, r]  (gk_proj[1] of
            // model/gla.py:96-97), rounded to the activation dtype like the GEMM it replaces
            float acc = b2 != nullptr ? to_f(b2[c]) : 0.f;
            for (int r = 0; r < R; ++r) acc = fmaf(to_f(lo[(size_t)b * ld_lo + r]), to_f(w2[(size_t)c * R + r]), acc);
            x = to_f(from_f<T>(acc));
        } else {
            x = to_f(gk_raw[(size_t)b * ldg + c]);
        }
        ef[(size_t)b * HK + c] = expf(logsigmoidf_(x) * inv_norm);
        return;
    }
    const T *x; const T *w; CT *cache; float *out; int D, ld; float mul = 1.f;
    if (c < HK) { x = xq; w = wq; cache = cq; out = qf; D = HK; ld = ld_qk; mul = scale; }
    else if (c < 2 * HK) { c -= HK; x = xk; w = wk; cache = ck; out = kf; D = HK; ld = ld_qk; }
    else { c -= 2 * HK; x = xv; w = wv; cache = cv; out = vf; D = HV; ld = ld_v; }
    float y = to_f(x[(size_t)b * ld + c]);
    if (w != nullptr) {
        // cache <- roll(cache, -1); cache[-1] = x; y = silu(sum_j cache[j] * w[j])   (convolution.py:197-204)
        CT *cp = cache + ((size_t)b * D + c) * W;
        const T *wp = w + (size_t)c * W;
        float acc = 0.f;
        for (int j = 0; j < W - 1; ++j) {
            const CT nxt = cp[j + 1];
            cp[j] = nxt;
            acc = fmaf(to_f(nxt), to_f(wp[j]), acc);
        }
        const CT last = from_f<CT>(y);
        cp[W - 1] = last;
        acc = fmaf(to_f(last), to_f(wp[W - 1]), acc);
        y = siluf_(acc);
        // the reference returns the conv output in the activation dtype before the GLA op
        y = to_f(from_f<T>(y));
    }
    out[(size_t)b * D + c] = y * mul;
}

template <int VEC> struct Vec;
template <> struct Vec<4> {  // fp32 state, 16 B
    static __device__ __forceinline__ void load(const float *p, float *x) {
        const float4 t = *reinterpret_cast<const float4 *>(p);
        x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
    }
    static __device__ __forceinline__ void store(float *p, const float *x) {
        *reinterpret_cast<float4 *>(p) = make_float4(x[0], x[1], x[2], x[3]);
    }
};
template <> struct Vec<8> {  // bf16 state, 16 B
    static __device__ __forceinline__ void load(const bf16 *p, float *x) {
        const uint4 t = *reinterpret_cast<const uint4 *>(p);
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); x[2 * i] = f.x; x[2 * i + 1] = f.y; }
    }
    static __device__ __forceinline__ void store(bf16 *p, const float *x) {
        uint4 t;
        __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
        *reinterpret_cast<uint4 *>(p) = t;
    }
};

constexpr int SW = 8;     // warps per CTA
// state rows in flight per thread: 4 when the grid oversubscribes the GPU (batch 128: 55 warps per SM), 8 at small batches, where
// the few resident warps cannot keep enough bytes in flight for the HBM rate (batch 32: 14 warps per SM, 16.3 us for 67 MB)

template <typename ST, int VEC, int SU>
__global__ void __launch_bounds__(SW * 32)
gla_step_state_kernel(ST *__restrict__ S, const float *__restrict__ qf, const float *__restrict__ kf,
                      const float *__restrict__ ef, const float *__restrict__ vf, float *__restrict__ of,
                      int K, int V) {
    __shared__ float sq[256], sk[256], se[256];
    __shared__ float red[SW][32 * VEC];
    const int bh = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int v0 = (blockIdx.x * 32 + lane) * VEC;
    const bool vok = v0 < V;          // V % VEC == 0 is checked by the host
    for (int i = tid; i < K; i += SW * 32) {
        sq[i] = qf[(size_t)bh * K + i]; sk[i] = kf[(size_t)bh * K + i]; se[i] = ef[(size_t)bh * K + i];
    }
    float vv[VEC], acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { vv[i] = vok ? vf[(size_t)bh * V + v0 + i] : 0.f; acc[i] = 0.f; }
    __syncthreads();
    ST *Sb = S + (size_t)bh * K * V + v0;
    if (vok) {
        int r = warp;
        for (; r + (SU - 1) * SW < K; r += SU * SW) {
            float s[SU][VEC];
#pragma unroll
            for (int u = 0; u < SU; ++u) Vec<VEC>::load(Sb + (size_t)(r + u * SW) * V, s[u]);
#pragma unroll
            for (int u = 0; u < SU; ++u) {
                const int rr = r + u * SW;
                const float e = se[rr], kk = sk[rr], qq = sq[rr];
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    s[u][i] = fmaf(s[u][i], e, kk * vv[i]);
                    acc[i] = fmaf(qq, s[u][i], acc[i]);
                }
                Vec<VEC>::store(Sb + (size_t)rr * V, s[u]);
            }
        }
        for (; r < K; r += SW) {
            float s[VEC];
            Vec<VEC>::load(Sb + (size_t)r * V, s);
            const float e = se[r], kk = sk[r], qq = sq[r];
#pragma unroll
            for (int i = 0; i < VEC; ++i) { s[i] = fmaf(s[i], e, kk * vv[i]); acc[i] = fmaf(qq, s[i], acc[i]); }
            Vec<VEC>::store(Sb + (size_t)r * V, s);
        }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) red[warp][lane * VEC + i] = acc[i];
    __syncthreads();
    for (int c = tid; c < 32 * VEC; c += SW * 32) {
        const int vc = blockIdx.x * 32 * VEC + c;
        if (vc < V) {
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < SW; ++w) sum += red[w][c];
            of[(size_t)bh * V + vc] = sum;
        }
    }
}

// RMSNorm over V of o (fp32 scratch) * w * swish(g); one CTA per (b,h) row, 4 consecutive columns per thread.
template <typename T>
__global__ void __launch_bounds__(256)
gla_step_norm_kernel(const float *__restrict__ of, const T *__restrict__ g, const T *__restrict__ w,
                     T *__restrict__ out, int V, float eps, int H, int ld_v) {
    __shared__ float red[8];
    const int row = blockIdx.x, tid = threadIdx.x;
    const int b = row / H, h = row - b * H;
    const float *x = of + (size_t)row * V;
    const T *gr = g + (size_t)b * ld_v + (size_t)h * V;
    T *orow = out + (size_t)row * V;
    float xv[4][4];
    float ss = 0.f;
    int nblk = 0;
    for (int c0 = tid * 4; c0 < V && nblk < 4; c0 += blockDim.x * 4, ++nblk) {
        const float4 t = *reinterpret_cast<const float4 *>(x + c0);
        // the reference's GLA op returns o in the activation dtype before the norm
        xv[nblk][0] = to_f(from_f<T>(t.x)); xv[nblk][1] = to_f(from_f<T>(t.y));
        xv[nblk][2] = to_f(from_f<T>(t.z)); xv[nblk][3] = to_f(from_f<T>(t.w));
#pragma unroll
        for (int i = 0; i < 4; ++i) ss = fmaf(xv[nblk][i], xv[nblk][i], ss);
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
    const float rstd = rsqrtf(tot / (float)V + eps);
    nblk = 0;
    for (int c0 = tid * 4; c0 < V && nblk < 4; c0 += blockDim.x * 4, ++nblk) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float gv = to_f(gr[c0 + i]);
            const float wv = w != nullptr ? to_f(w[c0 + i]) : 1.f;
            orow[c0 + i] = from_f<T>(xv[nblk][i] * rstd * wv * gv * sigmoidf_(gv));
        }
    }
}

template <typename T, typename CT>
int launch_step(const void *xq, const void *xk, const void *xv, const void *gk_raw, const void *g, const void *wq,
                const void *wk, const void *wv, void *cq, void *ck, void *cv, void *S, const void *norm_w,
                void *out, float *ws, int B, int H, int K, int V, int W, float scale, float gate_normalizer,
                float eps, int ld_qk, int ld_v, int ldg, cudaStream_t st, const void *lo = nullptr, int ld_lo = 0,
                const void *w2 = nullptr, const void *b2 = nullptr, int R = 0) {
    const int HK = H * K, HV = H * V;
    float *qf = ws, *kf = qf + (size_t)B * HK, *ef = kf + (size_t)B * HK, *vf = ef + (size_t)B * HK,
          *of = vf + (size_t)B * HV;
    const long long n1 = (long long)B * (3 * HK + HV);
    gla_step_prep_kernel<T, CT><<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(
        (const T *)xq, (const T *)xk, (const T *)xv, (const T *)gk_raw, (const T *)wq, (const T *)wk,
        (const T *)wv, (CT *)cq, (CT *)ck, (CT *)cv, qf, kf, ef, vf, B, HK, HV, W, scale, 1.f / gate_normalizer,
        ld_qk, ld_v, ldg, (const T *)lo, ld_lo, (const T *)w2, (const T *)b2, R);
    LINA_LAUNCH_OK("gla_step_prep_kernel");
    constexpr int VEC = sizeof(CT) == 4 ? 4 : 8;
    dim3 grid((V + 32 * VEC - 1) / (32 * VEC), B * H);
    static thread_local int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    }
    if ((long long)grid.x * grid.y <= 4LL * sms)
        gla_step_state_kernel<CT, VEC, 8><<<grid, SW * 32, 0, st>>>((CT *)S, qf, kf, ef, vf, of, K, V);
    else
        gla_step_state_kernel<CT, VEC, 4><<<grid, SW * 32, 0, st>>>((CT *)S, qf, kf, ef, vf, of, K, V);
    LINA_LAUNCH_OK("gla_step_state_kernel");
    const int rows = B * H;
    int nt = ((V / 4 + 31) / 32) * 32;           // V % 4 == 0 (checked); up to 4 float4 per thread -> V <= 4096
    if (nt > 256) nt = 256;
    gla_step_norm_kernel<T><<<rows, nt, 0, st>>>(of, (const T *)g, (const T *)norm_w, (T *)out, V, eps, H, ld_v);
    LINA_LAUNCH_OK("gla_step_norm_kernel");
    return LINA_OK;
}

}  // namespace

extern "C" size_t lina_gla_step_workspace_bytes(int B, int H, int K, int V) {
    return ((size_t)3 * B * H * K + (size_t)2 * B * H * V) * sizeof(float);
}

// ldx: common row stride (elements) of xq, xk, xv, g when they are column slices of ONE projection buffer
// (0 = each tensor dense); ldg: row stride of gk_raw (0 = dense).
static int gla_step_impl(const void *xq, const void *xk, const void *xv, const void *gk_raw, const void *g,
                         const void *wq, const void *wk, const void *wv, void *cq, void *ck, void *cv,
                         void *S, const void *norm_w, void *out, void *ws, int B, int H, int K, int V, int W,
                         int dtype, int state_dtype, float scale, float gate_normalizer, float eps,
                         int ldx, int ldg, void *stream, const void *lo, int ld_lo, const void *w2, const void *b2, int R) {
    LINA_REQUIRE(B > 0 && H > 0 && K > 0 && V > 0, LINA_ERR_BAD_ARG, "gla_step: non-positive size");
    LINA_REQUIRE(xq && xk && xv && (gk_raw || lo) && g && S && out && ws, LINA_ERR_BAD_ARG, "gla_step: null pointer");
    LINA_REQUIRE(lo == nullptr || (w2 != nullptr && R >= 1 && R <= 256 && ld_lo >= R), LINA_ERR_BAD_ARG,
                 "gla_step: low-rank gate needs w2 and 1 <= R <= 256");
    const bool conv = wq != nullptr;
    LINA_REQUIRE(!conv || (wk && wv && cq && ck && cv && W >= 1 && W <= 16), LINA_ERR_BAD_ARG,
                 "gla_step: short conv needs all three taps/states and 1 <= W <= 16");
    LINA_REQUIRE(K <= 256, LINA_ERR_UNSUPPORTED, "gla_step: K=%d > 256 not implemented", K);
    LINA_REQUIRE((long long)B * H <= 65535, LINA_ERR_UNSUPPORTED, "gla_step: B*H > 65535");
    LINA_REQUIRE(state_dtype == LINA_F32 || state_dtype == LINA_BF16, LINA_ERR_UNSUPPORTED,
                 "gla_step: state dtype must be f32 or bf16");
    LINA_REQUIRE(V % (state_dtype == LINA_F32 ? 4 : 8) == 0, LINA_ERR_UNSUPPORTED,
                 "gla_step: V=%d must be a multiple of %d", V, state_dtype == LINA_F32 ? 4 : 8);
    LINA_REQUIRE(gate_normalizer != 0.f, LINA_ERR_BAD_ARG, "gla_step: zero gate normalizer");
    LINA_REQUIRE(V <= 4096, LINA_ERR_UNSUPPORTED, "gla_step: V=%d > 4096 not implemented", V);
    LINA_REQUIRE(ldx == 0 || ldx >= H * V, LINA_ERR_BAD_ARG, "gla_step: ldx=%d smaller than a row", ldx);
    LINA_REQUIRE(ldg == 0 || ldg >= H * K, LINA_ERR_BAD_ARG, "gla_step: ldg=%d smaller than a row", ldg);
    const int ld_qk = ldx ? ldx : H * K, ld_v = ldx ? ldx : H * V, lg = ldg ? ldg : H * K;
    cudaStream_t st = (cudaStream_t)stream;
#define GO_(T, CT) return launch_step<T, CT>(xq, xk, xv, gk_raw, g, wq, wk, wv, cq, ck, cv, S, norm_w, out, \
                                             (float *)ws, B, H, K, V, W, scale, gate_normalizer, eps, ld_qk, ld_v, lg, st, \
                                             lo, ld_lo, w2, b2, R)
    if (dtype == LINA_F32 && state_dtype == LINA_F32) GO_(float, float);
    if (dtype == LINA_BF16 && state_dtype == LINA_BF16) GO_(bf16, bf16);
    if (dtype == LINA_BF16 && state_dtype == LINA_F32) GO_(bf16, float);
    if (dtype == LINA_F32 && state_dtype == LINA_BF16) GO_(float, bf16);
#undef GO_
    lina_set_error("gla_step: dtype %d / state dtype %d combination not implemented", dtype, state_dtype);
    return LINA_ERR_UNSUPPORTED;
}

extern "C" int lina_gla_step_ld(const void *xq, const void *xk, const void *xv, const void *gk_raw, const void *g,
                                const void *wq, const void *wk, const void *wv, void *cq, void *ck, void *cv,
                                void *S, const void *norm_w, void *out, void *ws, int B, int H, int K, int V, int W,
                                int dtype, int state_dtype, float scale, float gate_normalizer, float eps,
                                int ldx, int ldg, void *stream) {
    return gla_step_impl(xq, xk, xv, gk_raw, g, wq, wk, wv, cq, ck, cv, S, norm_w, out, ws, B, H, K, V, W, dtype, state_dtype,
                         scale, gate_normalizer, eps, ldx, ldg, stream, nullptr, 0, nullptr, nullptr, 0);
}

// Same with the gate logits computed in the kernel from the rank-R factors: lo [B, R] (row stride ld_lo; e.g. the last
// columns of the [q;k;v;g;gk0] projection), w2 [H*K, R], b2 [H*K] or NULL -- saves the gk_proj[1] GEMM launch per layer
// per token.
extern "C" int lina_gla_step_lr(const void *xq, const void *xk, const void *xv, const void *lo, int ld_lo, const void *w2,
                                const void *b2, int R, const void *g, const void *wq, const void *wk, const void *wv,
                                void *cq, void *ck, void *cv, void *S, const void *norm_w, void *out, void *ws, int B,
                                int H, int K, int V, int W, int dtype, int state_dtype, float scale,
                                float gate_normalizer, float eps, int ldx, void *stream) {
    return gla_step_impl(xq, xk, xv, nullptr, g, wq, wk, wv, cq, ck, cv, S, norm_w, out, ws, B, H, K, V, W, dtype, state_dtype,
                         scale, gate_normalizer, eps, ldx, 0, stream, lo, ld_lo, w2, b2, R);
}

extern "C" int lina_gla_step(const void *xq, const void *xk, const void *xv, const void *gk_raw, const void *g,
                             const void *wq, const void *wk, const void *wv, void *cq, void *ck, void *cv,
                             void *S, const void *norm_w, void *out, void *ws, int B, int H, int K, int V, int W,
                             int dtype, int state_dtype, float scale, float gate_normalizer, float eps,
                             void *stream) {
    return lina_gla_step_ld(xq, xk, xv, gk_raw, g, wq, wk, wv, cq, ck, cv, S, norm_w, out, ws, B, H, K, V, W, dtype,
                            state_dtype, scale, gate_normalizer, eps, 0, 0, stream);
}
