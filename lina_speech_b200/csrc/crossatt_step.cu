// One decode step of BlindCrossAttention's two single-head softmax attentions (model/crossatt.py:105-155, eval branch
// :13-19): per sequence  w = softmax(LN?(q) K^T / sqrt(d)),  out = w V  with K / V the memoised text-side tensors
// ([n, d] per sequence, or one shared [n, d] positional table).  The reference runs q@K^T, the scale, softmax and w@V as four
// launches (plus the LayerNorm) on [B,1,1,n] tensors; here one CTA per sequence does the lot: K and V rows stream once from
// L2 / HBM (16-byte loads), the scores live in shared memory.  Roundings follow the unfused bf16 sequence (scores, scaled
// scores, probabilities and output rounded to the activation dtype).
#include "common.cuh"

namespace {

constexpr int CA_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(CA_THREADS)
cross_att_step_kernel(const T *__restrict__ q_in, long long ldq, const T *__restrict__ ln_w, const T *__restrict__ ln_b, float eps,
                      const T *__restrict__ keys, long long key_bstride, const T *__restrict__ vals, long long val_bstride,
                      T *__restrict__ att_out, long long att_bstride, T *__restrict__ out, long long ldo, int n, int d, float scale) {
    extern __shared__ float sm[];
    float *qs = sm;                 // [d]   the (normalised) query
    float *sc = sm + d;             // [n]   scores -> probabilities
    __shared__ float red[CA_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const T *qr = q_in + (size_t)b * ldq;
    // ---- query (+ LayerNorm) ------------------------------------------------------------------------------------------
    float s = 0.f;
    for (int c = tid; c < d; c += CA_THREADS) { const float v = to_f(qr[c]); qs[c] = v; s += v; }
    if (ln_w != nullptr) {
        s = warp_sum(s);
        if (lane == 0) red[warp] = s;
        __syncthreads();
        float tot = 0.f;
        for (int w = 0; w < CA_THREADS / 32; ++w) tot += red[w];
        const float mean = tot / (float)d;
        __syncthreads();
        float ss = 0.f;
        for (int c = tid; c < d; c += CA_THREADS) { const float dv = qs[c] - mean; ss = fmaf(dv, dv, ss); }
        ss = warp_sum(ss);
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        float tv = 0.f;
        for (int w = 0; w < CA_THREADS / 32; ++w) tv += red[w];
        const float rstd = rsqrtf(tv / (float)d + eps);
        for (int c = tid; c < d; c += CA_THREADS)
            qs[c] = to_f(from_f<T>((qs[c] - mean) * rstd * to_f(ln_w[c]) + to_f(ln_b[c])));
    }
    __syncthreads();
    // ---- scores: a warp per key row, four rows' loads in flight ---------------------------------------------------------------
    const T *kb = keys + (size_t)b * key_bstride;
    constexpr int VEC = 16 / sizeof(T);
    constexpr int NW = CA_THREADS / 32;
    constexpr int MAXC = 4;                                       // chunks of VEC per lane held in registers (d <= 128 * VEC)
    const bool q_in_regs = d <= 32 * VEC * MAXC;
    float qr_[MAXC][VEC];
    if (q_in_regs) {                                              // lanes read qs with a stride of VEC floats: do it ONCE, not per key row
#pragma unroll
        for (int cc = 0; cc < MAXC; ++cc) {
            const int c = (cc * 32 + lane) * VEC;
#pragma unroll
            for (int i = 0; i < VEC; ++i) qr_[cc][i] = c < d ? qs[c + i] : 0.f;
        }
    }
    for (int j0 = warp; j0 < n; j0 += 4 * NW) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (q_in_regs) {
#pragma unroll
            for (int cc = 0; cc < MAXC; ++cc) {
                const int c = (cc * 32 + lane) * VEC;
                if (c < d) {
                    uint4 raw[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = j0 + u * NW;
                        raw[u] = j < n ? *reinterpret_cast<const uint4 *>(kb + (size_t)j * d + c) : make_uint4(0u, 0u, 0u, 0u);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const T *e = reinterpret_cast<const T *>(&raw[u]);
#pragma unroll
                        for (int i = 0; i < VEC; ++i) acc[u] = fmaf(qr_[cc][i], to_f(e[i]), acc[u]);
                    }
                }
            }
        } else {
            for (int c = lane * VEC; c < d; c += 32 * VEC) {
                uint4 raw[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + u * NW;
                    raw[u] = j < n ? *reinterpret_cast<const uint4 *>(kb + (size_t)j * d + c) : make_uint4(0u, 0u, 0u, 0u);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const T *e = reinterpret_cast<const T *>(&raw[u]);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[u] = fmaf(qs[c + i], to_f(e[i]), acc[u]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * NW;
            const float a = warp_sum(acc[u]);
            if (lane == 0 && j < n) sc[j] = to_f(from_f<T>(to_f(from_f<T>(a)) * scale));   // q@K^T rounded, then the scale rounded
        }
    }
    __syncthreads();
    // ---- softmax over the n keys ----------------------------------------------------------------------------------------------
    float m = -INFINITY;
    for (int j = tid; j < n; j += CA_THREADS) m = fmaxf(m, sc[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    float mx = red[0];
    for (int w = 1; w < CA_THREADS / 32; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float se = 0.f;
    for (int j = tid; j < n; j += CA_THREADS) { const float e = expf(sc[j] - mx); sc[j] = e; se += e; }
    se = warp_sum(se);
    if (lane == 0) red[warp] = se;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < CA_THREADS / 32; ++w) tot += red[w];
    const float inv = 1.f / tot;
    for (int j = tid; j < n; j += CA_THREADS) {
        const T p = from_f<T>(sc[j] * inv);
        sc[j] = to_f(p);
        if (att_out != nullptr) att_out[(size_t)b * att_bstride + j] = p;
    }
    __syncthreads();
    // ---- out = w V: a thread per VEC channels and slice of the keys, eight rows' loads in flight ----------------------------------
    const T *vb = vals + (size_t)b * val_bstride;
    const int groups = d / VEC;                                   // channel groups of VEC
    float *part = sm + d + n;                                     // [slices][d] partial sums (allocated by the host when slices > 1)
    if (groups <= CA_THREADS && CA_THREADS % groups == 0) {
        const int nslice = CA_THREADS / groups, slice = tid / groups, c = (tid - slice * groups) * VEC;
        const int per = (n + nslice - 1) / nslice, jb = slice * per, je = min(n, jb + per);
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
        for (int j0 = jb; j0 < je; j0 += 8) {
            uint4 raw[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                raw[u] = j0 + u < je ? *reinterpret_cast<const uint4 *>(vb + (size_t)(j0 + u) * d + c) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const T *e = reinterpret_cast<const T *>(&raw[u]);
                const float p = j0 + u < je ? sc[j0 + u] : 0.f;
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] = fmaf(p, to_f(e[i]), acc[i]);
            }
        }
        if (nslice > 1) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) part[(size_t)slice * d + c + i] = acc[i];
            __syncthreads();
            if (slice == 0) {
                for (int sl = 1; sl < nslice; ++sl)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[i] += part[(size_t)sl * d + c + i];
            }
        }
        if (slice == 0) {
            uint4 ov;
            T *oe = reinterpret_cast<T *>(&ov);
#pragma unroll
            for (int i = 0; i < VEC; ++i) oe[i] = from_f<T>(acc[i]);
            *reinterpret_cast<uint4 *>(out + (size_t)b * ldo + c) = ov;
        }
    } else {
        for (int c = tid * VEC; c < d; c += CA_THREADS * VEC) {
            float acc[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
            for (int j = 0; j < n; ++j) {
                const uint4 raw = *reinterpret_cast<const uint4 *>(vb + (size_t)j * d + c);
                const T *e = reinterpret_cast<const T *>(&raw);
                const float p = sc[j];
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] = fmaf(p, to_f(e[i]), acc[i]);
            }
            uint4 ov;
            T *oe = reinterpret_cast<T *>(&ov);
#pragma unroll
            for (int i = 0; i < VEC; ++i) oe[i] = from_f<T>(acc[i]);
            *reinterpret_cast<uint4 *>(out + (size_t)b * ldo + c) = ov;
        }
    }
}

}  // namespace

extern "C" int lina_cross_att_step(const void *q, long long ldq, const void *ln_w, const void *ln_b, float eps, const void *keys,
                                   long long key_bstride, const void *vals, long long val_bstride, void *att_out,
                                   long long att_bstride, void *out, long long ldo, int B, int n, int d, float scale, int dtype,
                                   void *stream) {
    LINA_REQUIRE(q && keys && vals && out, LINA_ERR_BAD_ARG, "cross_att_step: null pointer");
    LINA_REQUIRE(B > 0 && n > 0 && d > 0, LINA_ERR_BAD_ARG, "cross_att_step: bad size");
    LINA_REQUIRE(lina_dtype_ok(dtype), LINA_ERR_BAD_ARG, "cross_att_step: unknown dtype");
    LINA_REQUIRE((ln_w == nullptr) == (ln_b == nullptr), LINA_ERR_BAD_ARG, "cross_att_step: LayerNorm needs weight and bias");
    const int vec = 16 / (int)lina_dtype_size(dtype);
    LINA_REQUIRE(d % vec == 0 && ldo % vec == 0 && key_bstride % vec == 0 && val_bstride % vec == 0 &&
                     ((uintptr_t)keys & 15u) == 0 && ((uintptr_t)vals & 15u) == 0 && ((uintptr_t)out & 15u) == 0,
                 LINA_ERR_UNSUPPORTED, "cross_att_step: d and strides must be multiples of %d elements, tensors 16-byte aligned", vec);
    const int groups = d / vec, nslice = (groups <= CA_THREADS && CA_THREADS % groups == 0) ? CA_THREADS / groups : 1;
    const size_t smem = (size_t)(d + n + (nslice > 1 ? nslice * d : 0)) * sizeof(float);
    LINA_REQUIRE(smem <= 48 * 1024, LINA_ERR_UNSUPPORTED, "cross_att_step: d + n = %d too large", d + n);
    cudaStream_t st = (cudaStream_t)stream;
    LINA_DISPATCH_DTYPE(dtype, cross_att_step_kernel<T_><<<B, CA_THREADS, smem, st>>>(
                                   (const T_ *)q, ldq, (const T_ *)ln_w, (const T_ *)ln_b, eps, (const T_ *)keys, key_bstride,
                                   (const T_ *)vals, val_bstride, (T_ *)att_out, att_bstride, (T_ *)out, ldo, n, d, scale));
    LINA_LAUNCH_OK("cross_att_step_kernel");
    return LINA_OK;
}
