// GLA recurrence on CUDA cores: forward, backward (dq, dk, dv, dgk, dh0).
//
//   S_t = diag(exp(gk_t)) S_{t-1} + k_t^T v_t ;  o_t = scale * q_t S_t
//
// Spec: FLA/fla/ops/gla/naive.py:13-44; replaces the Triton kernels
// FLA/fla/ops/common/fused_recurrent.py:21-104 (fwd) and :118-257 (bwd).
//
// Layout of one CTA: a [K, 32] column tile of the state of one (batch, head) lives in
// registers -- warp w owns rows [w*KPW, (w+1)*KPW), lane l owns column v0 + l -- so the
// CTA owns every K row of its columns and the output reduction over K never leaves the
// CTA.  Time is serial; q/k/exp(gk) for TS steps are staged through shared memory as
// fp32 (exp taken once per element per CTA, broadcast-read as float4 by the warps).
// HBM traffic = the compulsory reads/writes (q,k,gk re-read once per column tile, from L2).
#include "common.cuh"

namespace {

constexpr int TS = 8;        // time steps staged per shared-memory round
constexpr int NWARP = 8;     // 256 threads
constexpr int BV = 32;       // state columns per CTA

template <int N, int STRIDE> struct TReduce {
    // halving butterfly: N values per lane -> N/2, exchanging with lane ^ STRIDE
    static __device__ __forceinline__ void run(float *p, int lane) {
        const bool hi = (lane & STRIDE) != 0;
#pragma unroll
        for (int j = 0; j < N / 2; ++j) {
            const float keep = hi ? p[j + N / 2] : p[j];
            const float send = hi ? p[j] : p[j + N / 2];
            p[j] = keep + __shfl_xor_sync(0xffffffffu, send, STRIDE);
        }
        TReduce<N / 2, STRIDE / 2>::run(p, lane);
    }
};
template <int STRIDE> struct TReduce<1, STRIDE> {
    static __device__ __forceinline__ void run(float *p, int) {
        float x = p[0];
#pragma unroll
        for (int s = STRIDE; s > 0; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
        p[0] = x;
    }
};
template <int N> struct TReduce<N, 0> {
    static __device__ __forceinline__ void run(float *, int) {}
};
template <> struct TReduce<1, 0> {
    static __device__ __forceinline__ void run(float *, int) {}
};

template <int KPW> struct Log2 { static constexpr int v = 1 + Log2<KPW / 2>::v; };
template <> struct Log2<1> { static constexpr int v = 0; };

// Sum p[0..KPW) over the 32 lanes; afterwards lane l holds (in the return value) the total of
// row (l >> (5 - log2 KPW)); lanes with the low (5 - log2 KPW) bits clear are the "owners".
template <int KPW> __device__ __forceinline__ float warp_transpose_reduce(float (&p)[KPW], int lane) {
    TReduce<KPW, 16>::run(p, lane);
    return p[0];
}

template <typename T, int KPW>
__global__ void __launch_bounds__(NWARP * 32)
gla_rec_fwd_kernel(const T *__restrict__ q, const T *__restrict__ k, const T *__restrict__ v,
                   const T *__restrict__ gk, const void *__restrict__ h0, int h0_dtype,
                   T *__restrict__ o, float *__restrict__ ht, int Tn, int K, int V, float scale,
                   const T *__restrict__ u, int H) {
    // u != nullptr selects the RWKV6 form (FLA/fla/ops/rwkv6/recurrent_naive.py:30-39): the output is taken
    // BEFORE the state update and the current token enters through the bonus u[h,k]:
    //     o_t = scale * q_t (S_{t-1} + diag(u) k_t^T v_t) ;  S_t = diag(exp(w_t)) S_{t-1} + k_t^T v_t
    constexpr int KP = NWARP * KPW;
    __shared__ __align__(16) float sq[TS][KP];
    __shared__ __align__(16) float sk[TS][KP];
    __shared__ __align__(16) float se[TS][KP];
    __shared__ float sv[TS][BV];
    __shared__ float so[TS][NWARP][BV];

    const int bh = blockIdx.y, v0 = blockIdx.x * BV;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int vcol = v0 + lane;
    const bool vok = vcol < V;
    const size_t qoff = (size_t)bh * Tn * K, voff = (size_t)bh * Tn * V, soff = (size_t)bh * K * V;
    const T *qb = q + qoff, *kb = k + qoff, *gb = gk + qoff, *vb = v + voff;
    T *ob = o + voff;

    float S[KPW], ub[KPW];
#pragma unroll
    for (int j = 0; j < KPW; ++j) {
        const int kk = warp * KPW + j;
        S[j] = (h0 != nullptr && kk < K && vok) ? load_dyn(h0, h0_dtype, soff + (size_t)kk * V + vcol) : 0.f;
        ub[j] = (u != nullptr && kk < K) ? to_f(u[(size_t)(bh % H) * K + kk]) : 0.f;
    }
    const bool rwkv = u != nullptr;

    for (int t0 = 0; t0 < Tn; t0 += TS) {
        for (int i = tid; i < TS * KP; i += NWARP * 32) {
            const int s = i / KP, kk = i - s * KP, t = t0 + s;
            float qv = 0.f, kv = 0.f, ev = 1.f;
            if (t < Tn && kk < K) {
                const size_t idx = (size_t)t * K + kk;
                qv = to_f(qb[idx]) * scale;
                kv = to_f(kb[idx]);
                ev = expf(to_f(gb[idx]));
            }
            sq[s][kk] = qv; sk[s][kk] = kv; se[s][kk] = ev;
        }
        {
            const int s = tid >> 5, t = t0 + s;
            sv[s][lane] = (t < Tn && vok) ? to_f(vb[(size_t)t * V + vcol]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < TS; ++s) {
            const float vv = sv[s][lane];
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < KPW; j += 4) {
                const int kk = warp * KPW + j;
                const float4 e4 = *reinterpret_cast<const float4 *>(&se[s][kk]);
                const float4 k4 = *reinterpret_cast<const float4 *>(&sk[s][kk]);
                const float4 q4 = *reinterpret_cast<const float4 *>(&sq[s][kk]);
                if (!rwkv) {
                    S[j + 0] = fmaf(S[j + 0], e4.x, k4.x * vv); acc = fmaf(q4.x, S[j + 0], acc);
                    S[j + 1] = fmaf(S[j + 1], e4.y, k4.y * vv); acc = fmaf(q4.y, S[j + 1], acc);
                    S[j + 2] = fmaf(S[j + 2], e4.z, k4.z * vv); acc = fmaf(q4.z, S[j + 2], acc);
                    S[j + 3] = fmaf(S[j + 3], e4.w, k4.w * vv); acc = fmaf(q4.w, S[j + 3], acc);
                } else {
                    const float kv0 = k4.x * vv, kv1 = k4.y * vv, kv2 = k4.z * vv, kv3 = k4.w * vv;
                    acc = fmaf(q4.x, fmaf(ub[j + 0], kv0, S[j + 0]), acc); S[j + 0] = fmaf(S[j + 0], e4.x, kv0);
                    acc = fmaf(q4.y, fmaf(ub[j + 1], kv1, S[j + 1]), acc); S[j + 1] = fmaf(S[j + 1], e4.y, kv1);
                    acc = fmaf(q4.z, fmaf(ub[j + 2], kv2, S[j + 2]), acc); S[j + 2] = fmaf(S[j + 2], e4.z, kv2);
                    acc = fmaf(q4.w, fmaf(ub[j + 3], kv3, S[j + 3]), acc); S[j + 3] = fmaf(S[j + 3], e4.w, kv3);
                }
            }
            so[s][warp][lane] = acc;
        }
        __syncthreads();
        {
            const int s = tid >> 5, t = t0 + s;
            if (t < Tn && vok) {
                float sum = 0.f;
#pragma unroll
                for (int w = 0; w < NWARP; ++w) sum += so[s][w][lane];
                ob[(size_t)t * V + vcol] = from_f<T>(sum);
            }
        }
        // next round's staging writes sq/sk/se/sv only after every warp passed the barrier above;
        // `so` is rewritten only after the next round's first barrier.
    }
    if (ht != nullptr && vok) {
#pragma unroll
        for (int j = 0; j < KPW; ++j) {
            const int kk = warp * KPW + j;
            if (kk < K) ht[soff + (size_t)kk * V + vcol] = S[j];
        }
    }
}

// ---- backward, sweep 1 (forward in time): recompute S_t, dq_t = scale * S_t do_t, and
//      c = sum_v dht * S_T (the final-state term of dgk).  Partial over this CTA's 32 columns,
//      accumulated across column tiles with fp32 atomics.
template <typename T, int KPW>
__global__ void __launch_bounds__(NWARP * 32)
gla_rec_bwd_dq_kernel(const T *__restrict__ k, const T *__restrict__ v, const T *__restrict__ gk,
                      const void *__restrict__ h0, int h0_dtype, const T *__restrict__ d_o,
                      const float *__restrict__ dht, float *__restrict__ dq_acc, float *__restrict__ cvec,
                      int Tn, int K, int V, float scale) {
    constexpr int KP = NWARP * KPW;
    constexpr int LG = Log2<KPW>::v;
    __shared__ __align__(16) float sk[TS][KP];
    __shared__ __align__(16) float se[TS][KP];
    __shared__ float sv[TS][BV];
    __shared__ float sdo[TS][BV];

    const int bh = blockIdx.y, v0 = blockIdx.x * BV;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int vcol = v0 + lane;
    const bool vok = vcol < V;
    const size_t qoff = (size_t)bh * Tn * K, voff = (size_t)bh * Tn * V, soff = (size_t)bh * K * V;
    const T *kb = k + qoff, *gb = gk + qoff, *vb = v + voff, *dob = d_o + voff;
    const bool owner = (lane & ((32 >> LG) - 1)) == 0;
    const int myrow = warp * KPW + (lane >> (5 - LG));

    float S[KPW];
#pragma unroll
    for (int j = 0; j < KPW; ++j) {
        const int kk = warp * KPW + j;
        S[j] = (h0 != nullptr && kk < K && vok) ? load_dyn(h0, h0_dtype, soff + (size_t)kk * V + vcol) : 0.f;
    }
    for (int t0 = 0; t0 < Tn; t0 += TS) {
        __syncthreads();
        for (int i = tid; i < TS * KP; i += NWARP * 32) {
            const int s = i / KP, kk = i - s * KP, t = t0 + s;
            float kv = 0.f, ev = 1.f;
            if (t < Tn && kk < K) {
                const size_t idx = (size_t)t * K + kk;
                kv = to_f(kb[idx]);
                ev = expf(to_f(gb[idx]));
            }
            sk[s][kk] = kv; se[s][kk] = ev;
        }
        {
            const int s = tid >> 5, t = t0 + s;
            const bool ok = t < Tn && vok;
            sv[s][lane] = ok ? to_f(vb[(size_t)t * V + vcol]) : 0.f;
            sdo[s][lane] = ok ? to_f(dob[(size_t)t * V + vcol]) : 0.f;
        }
        __syncthreads();
        for (int s = 0; s < TS; ++s) {
            const int t = t0 + s;
            if (t >= Tn) break;
            const float vv = sv[s][lane], dov = sdo[s][lane];
            float p[KPW];
#pragma unroll
            for (int j = 0; j < KPW; ++j) {
                const int kk = warp * KPW + j;
                S[j] = fmaf(S[j], se[s][kk], sk[s][kk] * vv);
                p[j] = S[j] * dov;
            }
            const float r = warp_transpose_reduce<KPW>(p, lane);
            if (owner && myrow < K) atomicAdd(&dq_acc[qoff + (size_t)t * K + myrow], scale * r);
        }
    }
    if (dht != nullptr) {
        float p[KPW];
#pragma unroll
        for (int j = 0; j < KPW; ++j) {
            const int kk = warp * KPW + j;
            p[j] = (kk < K && vok) ? S[j] * dht[soff + (size_t)kk * V + vcol] : 0.f;
        }
        const float r = warp_transpose_reduce<KPW>(p, lane);
        if (owner && myrow < K) atomicAdd(&cvec[(size_t)bh * K + myrow], r);
    }
}

// ---- backward, sweep 2 (reverse in time): dS carried in registers.
//      dS_t = e^{g_{t+1}} dS_{t+1} + scale q_t do_t^T ; dk_t = dS_t v_t ; dv_t = dS_t^T k_t ; dh0 = e^{g_0} dS_0
template <typename T, int KPW>
__global__ void __launch_bounds__(NWARP * 32)
gla_rec_bwd_dkv_kernel(const T *__restrict__ q, const T *__restrict__ k, const T *__restrict__ v,
                       const T *__restrict__ gk, const T *__restrict__ d_o, const float *__restrict__ dht,
                       float *__restrict__ dk_acc, T *__restrict__ dv, float *__restrict__ dh0,
                       int Tn, int K, int V, float scale) {
    constexpr int KP = NWARP * KPW;
    constexpr int LG = Log2<KPW>::v;
    __shared__ __align__(16) float sq[TS][KP];
    __shared__ __align__(16) float sk[TS][KP];
    __shared__ __align__(16) float se[TS][KP];
    __shared__ float sv[TS][BV];
    __shared__ float sdo[TS][BV];
    __shared__ float so[TS][NWARP][BV];

    const int bh = blockIdx.y, v0 = blockIdx.x * BV;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int vcol = v0 + lane;
    const bool vok = vcol < V;
    const size_t qoff = (size_t)bh * Tn * K, voff = (size_t)bh * Tn * V, soff = (size_t)bh * K * V;
    const T *qb = q + qoff, *kb = k + qoff, *gb = gk + qoff, *vb = v + voff, *dob = d_o + voff;
    T *dvb = dv + voff;
    const bool owner = (lane & ((32 >> LG) - 1)) == 0;
    const int myrow = warp * KPW + (lane >> (5 - LG));

    float dS[KPW];
#pragma unroll
    for (int j = 0; j < KPW; ++j) {
        const int kk = warp * KPW + j;
        dS[j] = (dht != nullptr && kk < K && vok) ? dht[soff + (size_t)kk * V + vcol] : 0.f;
    }
    const int nstage = (Tn + TS - 1) / TS;
    for (int st = nstage - 1; st >= 0; --st) {
        const int t0 = st * TS;
        for (int i = tid; i < TS * KP; i += NWARP * 32) {
            const int s = i / KP, kk = i - s * KP, t = t0 + s;
            float qv = 0.f, kv = 0.f, ev = 1.f;
            if (t < Tn && kk < K) {
                const size_t idx = (size_t)t * K + kk;
                qv = to_f(qb[idx]) * scale;
                kv = to_f(kb[idx]);
                ev = expf(to_f(gb[idx]));
            }
            sq[s][kk] = qv; sk[s][kk] = kv; se[s][kk] = ev;
        }
        {
            const int s = tid >> 5, t = t0 + s;
            const bool ok = t < Tn && vok;
            sv[s][lane] = ok ? to_f(vb[(size_t)t * V + vcol]) : 0.f;
            sdo[s][lane] = ok ? to_f(dob[(size_t)t * V + vcol]) : 0.f;
        }
        __syncthreads();
        for (int s = TS - 1; s >= 0; --s) {
            const int t = t0 + s;
            if (t >= Tn) continue;
            const float vv = sv[s][lane], dov = sdo[s][lane];
            float p[KPW];
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < KPW; ++j) {
                const int kk = warp * KPW + j;
                dS[j] = fmaf(sq[s][kk], dov, dS[j]);
                p[j] = dS[j] * vv;
                acc = fmaf(dS[j], sk[s][kk], acc);
            }
            so[s][warp][lane] = acc;
            const float r = warp_transpose_reduce<KPW>(p, lane);
            if (owner && myrow < K) atomicAdd(&dk_acc[qoff + (size_t)t * K + myrow], r);
#pragma unroll
            for (int j = 0; j < KPW; ++j) dS[j] *= se[s][warp * KPW + j];
        }
        __syncthreads();
        {
            const int s = tid >> 5, t = t0 + s;
            if (t < Tn && vok) {
                float sum = 0.f;
#pragma unroll
                for (int w = 0; w < NWARP; ++w) sum += so[s][w][lane];
                dvb[(size_t)t * V + vcol] = from_f<T>(sum);
            }
        }
        __syncthreads();
    }
    if (dh0 != nullptr && vok) {
#pragma unroll
        for (int j = 0; j < KPW; ++j) {
            const int kk = warp * KPW + j;
            if (kk < K) dh0[soff + (size_t)kk * V + vcol] = dS[j];
        }
    }
}

// ---- backward, finalize: dgk = reversed cumsum_t(dq*q - dk*k) + c ; cast dq, dk to the I/O dtype.
//      (FLA/fla/ops/common/fused_recurrent.py:335-342, plus the dht term the reference drops.)
template <typename T>
__global__ void __launch_bounds__(128)
gla_rec_bwd_finalize_kernel(const T *__restrict__ q, const T *__restrict__ k,
                            const float *__restrict__ dq_acc, const float *__restrict__ dk_acc,
                            const float *__restrict__ cvec, T *__restrict__ dq, T *__restrict__ dk,
                            T *__restrict__ dgk, int Tn, int K) {
    const int bh = blockIdx.y, kk = blockIdx.x * 128 + threadIdx.x;
    if (kk >= K) return;
    const size_t base = (size_t)bh * Tn * K + kk;
    float run = cvec[(size_t)bh * K + kk];
    for (int t = Tn - 1; t >= 0; --t) {
        const size_t idx = base + (size_t)t * K;
        const float a = dq_acc[idx], b = dk_acc[idx];
        run += a * to_f(q[idx]) - b * to_f(k[idx]);
        dgk[idx] = from_f<T>(run);
        dq[idx] = from_f<T>(a);
        dk[idx] = from_f<T>(b);
    }
}

template <typename T>
int launch_fwd(const void *q, const void *k, const void *v, const void *gk, const void *h0, int h0_dtype,
               void *o, float *ht, int B, int H, int Tn, int K, int V, float scale, cudaStream_t st,
               const void *u = nullptr) {
    dim3 grid((V + BV - 1) / BV, B * H), block(NWARP * 32);
#define L_(KPW) gla_rec_fwd_kernel<T, KPW><<<grid, block, 0, st>>>((const T *)q, (const T *)k, (const T *)v, \
        (const T *)gk, h0, h0_dtype, (T *)o, ht, Tn, K, V, scale, (const T *)u, H)
    if (K <= 32) L_(4); else if (K <= 64) L_(8); else if (K <= 128) L_(16); else L_(32);
#undef L_
    LINA_LAUNCH_OK("gla_rec_fwd_kernel");
    return LINA_OK;
}

template <typename T>
int launch_bwd(const void *q, const void *k, const void *v, const void *gk, const void *h0, int h0_dtype,
               const void *d_o, const float *dht, void *dq, void *dk, void *dv, void *dgk, float *dh0,
               float *dq_acc, float *dk_acc, float *cvec, int B, int H, int Tn, int K, int V, float scale,
               cudaStream_t st) {
    dim3 grid((V + BV - 1) / BV, B * H), block(NWARP * 32);
#define L1_(KPW) gla_rec_bwd_dq_kernel<T, KPW><<<grid, block, 0, st>>>((const T *)k, (const T *)v, (const T *)gk, \
        h0, h0_dtype, (const T *)d_o, dht, dq_acc, cvec, Tn, K, V, scale)
#define L2_(KPW) gla_rec_bwd_dkv_kernel<T, KPW><<<grid, block, 0, st>>>((const T *)q, (const T *)k, (const T *)v, \
        (const T *)gk, (const T *)d_o, dht, dk_acc, (T *)dv, dh0, Tn, K, V, scale)
    if (K <= 32) { L1_(4); L2_(4); } else if (K <= 64) { L1_(8); L2_(8); }
    else if (K <= 128) { L1_(16); L2_(16); } else { L1_(32); L2_(32); }
#undef L1_
#undef L2_
    LINA_LAUNCH_OK("gla_rec_bwd_{dq,dkv}_kernel");
    dim3 g3((K + 127) / 128, B * H);
    gla_rec_bwd_finalize_kernel<T><<<g3, 128, 0, st>>>((const T *)q, (const T *)k, dq_acc, dk_acc, cvec,
                                                       (T *)dq, (T *)dk, (T *)dgk, Tn, K);
    LINA_LAUNCH_OK("gla_rec_bwd_finalize_kernel");
    return LINA_OK;
}

int check_dims(int B, int H, int T, int K, int V, int dtype) {
    LINA_REQUIRE(B > 0 && H > 0 && T > 0 && K > 0 && V > 0, LINA_ERR_BAD_ARG,
                 "gla: non-positive size B=%d H=%d T=%d K=%d V=%d", B, H, T, K, V);
    LINA_REQUIRE(lina_dtype_ok(dtype), LINA_ERR_BAD_ARG, "gla: unknown dtype %d", dtype);
    LINA_REQUIRE(K <= 256, LINA_ERR_UNSUPPORTED, "gla: head key dim K=%d > 256 not implemented", K);
    LINA_REQUIRE((long long)B * H <= 65535, LINA_ERR_UNSUPPORTED, "gla: B*H=%lld > 65535", (long long)B * H);
    return LINA_OK;
}

}  // namespace

int lina_gla_recurrent_fwd_impl(const void *q, const void *k, const void *v, const void *gk, const void *h0,
                                int h0_dtype, void *o, float *ht, int B, int H, int T, int K, int V, int dtype,
                                float scale, void *stream) {
    int rc = check_dims(B, H, T, K, V, dtype);
    if (rc) return rc;
    LINA_REQUIRE(q && k && v && gk && o, LINA_ERR_BAD_ARG, "gla_recurrent_fwd: null tensor pointer");
    LINA_REQUIRE(h0 == nullptr || lina_dtype_ok(h0_dtype), LINA_ERR_BAD_ARG, "gla_recurrent_fwd: bad h0 dtype");
    LINA_DISPATCH_DTYPE(dtype, return launch_fwd<T_>(q, k, v, gk, h0, h0_dtype, o, ht, B, H, T, K, V, scale,
                                                      (cudaStream_t)stream));
    return LINA_OK;
}

extern "C" int lina_gla_recurrent_fwd(const void *q, const void *k, const void *v, const void *gk,
                                      const void *h0, int h0_dtype, void *o, float *ht, int B, int H, int T,
                                      int K, int V, int dtype, float scale, void *stream) {
    return lina_gla_recurrent_fwd_impl(q, k, v, gk, h0, h0_dtype, o, ht, B, H, T, K, V, dtype, scale, stream);
}

// RWKV6 recurrence, forward (FLA/fla/ops/rwkv6/recurrent_fuse.py:335-368 -> kernel :16-82; spec recurrent_naive.py:8-42)
extern "C" int lina_rwkv6_recurrent_fwd(const void *r, const void *k, const void *v, const void *w, const void *u,
                                        const void *h0, int h0_dtype, void *o, float *ht, int B, int H, int T, int K,
                                        int V, int dtype, float scale, void *stream) {
    int rc = check_dims(B, H, T, K, V, dtype);
    if (rc) return rc;
    LINA_REQUIRE(r && k && v && w && u && o, LINA_ERR_BAD_ARG, "rwkv6_recurrent_fwd: null tensor pointer");
    LINA_REQUIRE(h0 == nullptr || lina_dtype_ok(h0_dtype), LINA_ERR_BAD_ARG, "rwkv6_recurrent_fwd: bad h0 dtype");
    LINA_DISPATCH_DTYPE(dtype, return launch_fwd<T_>(r, k, v, w, h0, h0_dtype, o, ht, B, H, T, K, V, scale,
                                                      (cudaStream_t)stream, u));
    return LINA_OK;
}

extern "C" size_t lina_gla_recurrent_bwd_workspace_bytes(int B, int H, int T, int K, int V) {
    (void)V;
    return ((size_t)2 * B * H * T * K + (size_t)B * H * K) * sizeof(float);
}

extern "C" int lina_gla_recurrent_bwd(const void *q, const void *k, const void *v, const void *gk,
                                      const void *h0, int h0_dtype, const void *d_o, const float *dht,
                                      void *dq, void *dk, void *dv, void *dgk, float *dh0, void *ws,
                                      int B, int H, int T, int K, int V, int dtype, float scale, void *stream) {
    int rc = check_dims(B, H, T, K, V, dtype);
    if (rc) return rc;
    LINA_REQUIRE(q && k && v && gk && d_o && dq && dk && dv && dgk && ws, LINA_ERR_BAD_ARG,
                 "gla_recurrent_bwd: null tensor pointer");
    LINA_REQUIRE(h0 == nullptr || lina_dtype_ok(h0_dtype), LINA_ERR_BAD_ARG, "gla_recurrent_bwd: bad h0 dtype");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)B * H * T * K;
    float *dq_acc = (float *)ws, *dk_acc = dq_acc + n, *cvec = dk_acc + n;
    LINA_CUDA_OK(cudaMemsetAsync(ws, 0, lina_gla_recurrent_bwd_workspace_bytes(B, H, T, K, V), st));
    LINA_DISPATCH_DTYPE(dtype, return launch_bwd<T_>(q, k, v, gk, h0, h0_dtype, d_o, dht, dq, dk, dv, dgk, dh0,
                                                      dq_acc, dk_acc, cvec, B, H, T, K, V, scale, st));
    return LINA_OK;
}
