// Post-projection pass of a GLA mixer over a whole sequence, ONE launch:
//   q, k, v <- SiLU(depthwise causal conv_4(x_q | x_k | x_v))      (FLA/fla/modules/convolution.py:141-178 -> causal-conv1d)
//   gk      <- logsigmoid(gk_raw) / normalizer [clamped]            (model/gla.py:174-181)
// The three conv inputs may be column slices of one [q;k;v;g] projection buffer (row stride `ldx`), so the
// concatenated GEMM needs no split copies.  Pure HBM streaming: every element is read once (+3 halo rows per
// TL-row tile, L1/L2 hits) and written once.
//
// Work decomposition: one thread = 16 bytes of channels x TL consecutive time steps of one sequence.  All TL+3
// 16-byte loads of the thread are issued before the first use (>= 176 B in flight per thread): the previous
// kernel (4 loads in flight, 88 registers) reached 32 % of the measured HBM peak.
#include "common.cuh"
#include "sm100.cuh"
#include "packed.cuh"

int g_lina_variant[16] = {0};

extern "C" int lina_debug_set_variant(int key, int value) {
    if (key < 0 || key >= 16) return LINA_ERR_BAD_ARG;
    g_lina_variant[key] = value;
    return LINA_OK;
}

namespace {

using sm100::ex2_approx;
using sm100::pack_bf16;

constexpr int PREP_THREADS = 128;

struct PrepSeg {
    const void *x;      // [B, L, ldx] rows (conv: raw projection; gate: gate logits)
    const void *w;      // conv taps [D, 4] (dtype) or nullptr for a gate segment
    void *y;            // [B, L, D] contiguous
    void *cache;        // conv state [B, D, 4] or nullptr
    long long ldx;      // row stride of x in elements
    int D;
    int kind;           // 0 = conv(+silu), 1 = gate
};
struct PrepArgs {
    PrepSeg seg[4];
    int B, L, cache_dtype, silu, use_clamp;
    float inv_norm, clamp_min;
};

__device__ __forceinline__ float logsigmoid_fast_(float x) { return fminf(x, 0.f) - __logf(1.f + __expf(-fabsf(x))); }

template <typename T> __device__ __forceinline__ void unpack16(const uint4 &raw, float *f) {
    constexpr int VEC = 16 / sizeof(T);
    const T *e = reinterpret_cast<const T *>(&raw);
#pragma unroll
    for (int c = 0; c < VEC; ++c) f[c] = to_f(e[c]);
}

template <typename T, int TL>
__global__ void __launch_bounds__(PREP_THREADS, TL == 8 ? 4 : 2)
gla_prep_kernel(const __grid_constant__ PrepArgs a) {
    constexpr int VEC = 16 / sizeof(T);
    const PrepSeg &sg = a.seg[blockIdx.y];
    const int L = a.L;
    const int nv = sg.D / VEC;
    const int tiles = (L + TL - 1) / TL;
    const long long idx = (long long)blockIdx.x * PREP_THREADS + threadIdx.x;
    const int dv = (int)(idx % nv);
    const long long rest = idx / nv;
    const int tile = (int)(rest % tiles);
    const long long b = rest / tiles;
    if (b >= a.B) return;
    const int d0 = dv * VEC, l0 = tile * TL;
    const T *xb = reinterpret_cast<const T *>(sg.x) + (size_t)b * L * sg.ldx + d0;
    T *yb = reinterpret_cast<T *>(sg.y) + (size_t)b * L * sg.D + d0;

    if (sg.kind == 1) {
        // ---- gate: the reference rounds logsigmoid to the activation dtype before the division ----
        uint4 raw[TL];
#pragma unroll
        for (int i = 0; i < TL; ++i) {
            const int l = l0 + i;
            raw[i] = l < L ? *reinterpret_cast<const uint4 *>(xb + (size_t)l * sg.ldx) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int i = 0; i < TL; ++i) {
            const int l = l0 + i;
            if (l < L) {
                float f[VEC];
                unpack16<T>(raw[i], f);
                uint4 outr;
                T *oe = reinterpret_cast<T *>(&outr);
#pragma unroll
                for (int c = 0; c < VEC; ++c) {
                    float g = to_f(from_f<T>(sizeof(T) == 4 ? logsigmoidf_(f[c]) : logsigmoid_fast_(f[c]))) * a.inv_norm;
                    if (a.use_clamp) g = fmaxf(g, a.clamp_min);
                    oe[c] = from_f<T>(g);
                }
                *reinterpret_cast<uint4 *>(yb + (size_t)l * sg.D) = outr;
            }
        }
        return;
    }

    // ---- depthwise causal conv, 4 taps, optional SiLU ----
    uint4 raw[TL + 3];
#pragma unroll
    for (int i = 0; i < TL + 3; ++i) {
        const int l = l0 - 3 + i;
        raw[i] = (l >= 0 && l < L) ? *reinterpret_cast<const uint4 *>(xb + (size_t)l * sg.ldx) : make_uint4(0, 0, 0, 0);
    }
    // cache[b, d, j] = x[L-4+j] (zero left-padded when L < 4), written by the tile that holds the last row
    if (sg.cache != nullptr && l0 + TL >= L) {
#pragma unroll
        for (int i = 0; i < TL + 3; ++i) {
            const int l = l0 - 3 + i;
            const int j = l - (L - 4);
            if (j >= 0 && j < 4) {
                float f[VEC];
                unpack16<T>(raw[i], f);                   // rows < 0 were loaded as zeros
#pragma unroll
                for (int c = 0; c < VEC; ++c)
                    store_dyn(sg.cache, a.cache_dtype, ((size_t)b * sg.D + d0 + c) * 4 + j, f[c]);
            }
        }
    }
    // two passes over the rows, one per half of the thread's channels: only HV channels' taps and window are live
    constexpr int HV = VEC / 2;
    constexpr int HW = 2;                                 // 32-bit words per half of a 16-byte vector
    uint32_t outw[TL][4];
    const T *wp = reinterpret_cast<const T *>(sg.w) + (size_t)d0 * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float wt[HV][4];                                  // taps w[(d0 + h*HV + c)*4 + j]: HV*4 elements
        {
            constexpr int NV16 = HV * 4 / VEC;            // 16-byte vectors holding them (2 for 16-bit, 2 for fp32)
#pragma unroll
            for (int r = 0; r < NV16; ++r) {
                float f[VEC];
                unpack16<T>(*reinterpret_cast<const uint4 *>(wp + h * HV * 4 + r * VEC), f);
#pragma unroll
                for (int e = 0; e < VEC; ++e) wt[(r * VEC + e) >> 2][(r * VEC + e) & 3] = f[e];
            }
        }
        auto half = [&](const uint4 &v, float *f) {       // channels [h*HV, (h+1)*HV) of a 16-byte vector
            const uint32_t w2[HW] = {h == 0 ? v.x : v.z, h == 0 ? v.y : v.w};
            const T *e = reinterpret_cast<const T *>(w2);
#pragma unroll
            for (int c = 0; c < HV; ++c) f[c] = to_f(e[c]);
        };
        float win[3][HV];
        half(raw[0], win[0]);
        half(raw[1], win[1]);
        half(raw[2], win[2]);
#pragma unroll
        for (int i = 0; i < TL; ++i) {
            float xv[HV];
            half(raw[i + 3], xv);
            uint32_t ow[HW];
            T *oe = reinterpret_cast<T *>(ow);
#pragma unroll
            for (int c = 0; c < HV; ++c) {
                float acc = win[0][c] * wt[c][0];
                acc = fmaf(win[1][c], wt[c][1], acc);
                acc = fmaf(win[2][c], wt[c][2], acc);
                acc = fmaf(xv[c], wt[c][3], acc);
                oe[c] = from_f<T>(a.silu ? acc * sigmoid_io<T>(acc) : acc);
                win[0][c] = win[1][c]; win[1][c] = win[2][c]; win[2][c] = xv[c];
            }
            outw[i][2 * h] = ow[0];
            outw[i][2 * h + 1] = ow[1];
        }
    }
#pragma unroll
    for (int i = 0; i < TL; ++i) {
        const int l = l0 + i;
        if (l < L) *reinterpret_cast<uint4 *>(yb + (size_t)l * sg.D) = make_uint4(outw[i][0], outw[i][1], outw[i][2], outw[i][3]);
    }
}

// bf16 depthwise conv_4 + SiLU, packed math: thread = 16 bytes of channels x TL rows, all loads up front.
template <int TL>
__global__ void __launch_bounds__(PREP_THREADS, TL == 8 ? 4 : 2)
conv4_silu_bf16_kernel(const bf16 *__restrict__ x, long long ldx, const bf16 *__restrict__ w, bf16 *__restrict__ y,
                       void *__restrict__ cache, int cache_dtype, int B, int L, int D, int silu) {
    const int nv = D / 8;
    const int tiles = (L + TL - 1) / TL;
    const long long idx = (long long)blockIdx.x * PREP_THREADS + threadIdx.x;
    const int dv = (int)(idx % nv);
    const long long rest = idx / nv;
    const int tile = (int)(rest % tiles);
    const long long b = rest / tiles;
    if (b >= B) return;
    const int d0 = dv * 8, l0 = tile * TL;
    const bf16 *xb = x + (size_t)b * L * ldx + d0;
    bf16 *yb = y + (size_t)b * L * D + d0;
    uint4 raw[TL + 3];
#pragma unroll
    for (int i = 0; i < TL + 3; ++i) {
        const int l = l0 - 3 + i;
        raw[i] = (l >= 0 && l < L) ? *reinterpret_cast<const uint4 *>(xb + (size_t)l * ldx) : make_uint4(0, 0, 0, 0);
    }
    if (cache != nullptr && l0 + TL >= L) {               // cache[b, d, j] = x[L-4+j], zero left-padded
#pragma unroll
        for (int i = 0; i < TL + 3; ++i) {
            const int l = l0 - 3 + i;
            const int j = l - (L - 4);
            if (j >= 0 && j < 4) {
                const uint32_t wds[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float2 f = bf2_to_f2(wds[p]);
                    store_dyn(cache, cache_dtype, ((size_t)b * D + d0 + 2 * p) * 4 + j, f.x);
                    store_dyn(cache, cache_dtype, ((size_t)b * D + d0 + 2 * p + 1) * 4 + j, f.y);
                }
            }
        }
    }
    uint32_t outw[TL][4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {                         // channel pair by channel pair: 4 tap pairs + 3 window pairs live
        float2 wt[4];
        load_taps2(w + (size_t)(d0 + 2 * p) * 4, wt);
        auto word = [&](const uint4 &v) { return p == 0 ? v.x : (p == 1 ? v.y : (p == 2 ? v.z : v.w)); };
        float2 win[3] = {bf2_to_f2(word(raw[0])), bf2_to_f2(word(raw[1])), bf2_to_f2(word(raw[2]))};
#pragma unroll
        for (int i = 0; i < TL; ++i) {
            float2 acc = conv4_2(win, bf2_to_f2(word(raw[i + 3])), wt);
            if (silu) acc = silu2(acc);
            outw[i][p] = pack_bf16(acc.x, acc.y);
        }
    }
#pragma unroll
    for (int i = 0; i < TL; ++i) {
        const int l = l0 + i;
        if (l < L) *reinterpret_cast<uint4 *>(yb + (size_t)l * D) = make_uint4(outw[i][0], outw[i][1], outw[i][2], outw[i][3]);
    }
}

// ------------------------------------------------------------------------------------------------
// q/k side of the post-projection pass with the CHUNK GATING of the tensor-core GLA kernel folded in:
//   q = SiLU(conv_4(x_q)), k = SiLU(conv_4(x_k)), gk = bf16(logsigmoid(g_raw)) / normalizer      (as above)
//   G_t = sum_{s <= t, s in the 64-token chunk} gk_s        (fp32, kept in log2 units)
//   q~_t = scale * q_t * e^{G_t}   k~_t = k_t * e^{-G_t}   (bf16, the MMA operands)   decay[b,h,n,:] = e^{G_C}
// so the GLA kernel needs no gate pre-pass (it ran 4x redundantly, once per V slice, and was MUFU/issue bound)
// and never reads gk.  One thread = 4 channels x one whole chunk (64 rows, serial: the cumsum is thread-local).
struct GateArgs {
    const bf16 *xq, *xk, *graw, *wq, *wk;
    bf16 *qg, *kg;
    float *decay;
    void *cq, *ck;
    long long ldq, ldk, ldg;
    int B, L, Dk, K, NT, cache_dtype;
    float log2_scale, gate_c;        // gate_c = log2(e) / normalizer
    int *envelope_flag;              // set to 1 when a chunk's summed log2 gate leaves the single-pivot range (nullable)
};

// the tensor-core kernel keeps one pivot per 64-token chunk: k~ = k e^-G needs |G| (natural log) below ~85; flag at 80
constexpr float GATE_ENVELOPE_LOG2 = -80.f * 1.44269504088896340736f;

constexpr int GC = 64;               // chunk length of gla_chunk_sm100.cu
constexpr int GRB = 4;               // rows per load batch

__device__ __forceinline__ float logsigmoid_bf16r(float x) {   // bf16-rounded like the reference's activation dtype
    return __bfloat162float(__float2bfloat16_rn(logsigmoid_fast_(x)));
}

__global__ void __launch_bounds__(PREP_THREADS, 4)
qk_gate_bf16_kernel(const __grid_constant__ GateArgs a) {
    const int ng = a.Dk / 4;
    const long long idx = (long long)blockIdx.x * PREP_THREADS + threadIdx.x;
    const int cg = (int)(idx % ng);
    const long long rest = idx / ng;
    const int n = (int)(rest % a.NT);
    const long long b = rest / a.NT;
    if (b >= a.B) return;
    const int L = a.L, d0 = cg * 4, t0 = n * GC;
    const bf16 *xq = a.xq + (size_t)b * L * a.ldq + d0, *xk = a.xk + (size_t)b * L * a.ldk + d0;
    const bf16 *gr = a.graw + (size_t)b * L * a.ldg + d0;
    bf16 *qo = a.qg + (size_t)b * L * a.Dk + d0, *ko = a.kg + (size_t)b * L * a.Dk + d0;
    float2 wq[2][4], wk[2][4];
    load_taps2(a.wq + (size_t)d0 * 4, wq[0]); load_taps2(a.wq + (size_t)(d0 + 2) * 4, wq[1]);
    load_taps2(a.wk + (size_t)d0 * 4, wk[0]); load_taps2(a.wk + (size_t)(d0 + 2) * 4, wk[1]);
    float2 winq[2][3], wink[2][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int l = t0 - 3 + i;
        uint2 rq = make_uint2(0, 0), rk = make_uint2(0, 0);
        if (l >= 0) {
            rq = *reinterpret_cast<const uint2 *>(xq + (size_t)l * a.ldq);
            rk = *reinterpret_cast<const uint2 *>(xk + (size_t)l * a.ldk);
        }
        winq[0][i] = bf2_to_f2(rq.x); winq[1][i] = bf2_to_f2(rq.y);
        wink[0][i] = bf2_to_f2(rk.x); wink[1][i] = bf2_to_f2(rk.y);
    }
    float2 G[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};          // log2 units
    const int nrow = min(GC, L - t0);
#pragma unroll 1
    for (int r0 = 0; r0 < nrow; r0 += GRB) {
        uint2 rq[GRB], rk[GRB], rg[GRB];
#pragma unroll
        for (int i = 0; i < GRB; ++i) {
            const int l = t0 + r0 + i;
            if (r0 + i < nrow) {
                rq[i] = *reinterpret_cast<const uint2 *>(xq + (size_t)l * a.ldq);
                rk[i] = *reinterpret_cast<const uint2 *>(xk + (size_t)l * a.ldk);
                rg[i] = *reinterpret_cast<const uint2 *>(gr + (size_t)l * a.ldg);
            } else {
                rq[i] = rk[i] = rg[i] = make_uint2(0, 0);
            }
        }
#pragma unroll
        for (int i = 0; i < GRB; ++i) {
            if (r0 + i >= nrow) break;                    // block-uniform: rows past the sequence end are not gated
            const int l = t0 + r0 + i;
            uint32_t oq[2], ok[2];
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const uint32_t wq_ = p == 0 ? rq[i].x : rq[i].y, wk_ = p == 0 ? rk[i].x : rk[i].y;
                const uint32_t wg_ = p == 0 ? rg[i].x : rg[i].y;
                const float2 qv = silu2(conv4_2(winq[p], bf2_to_f2(wq_), wq[p]));
                const float2 kv = silu2(conv4_2(wink[p], bf2_to_f2(wk_), wk[p]));
                const float2 gx = bf2_to_f2(wg_);
                const float2 ls = make_float2(logsigmoid_bf16r(gx.x), logsigmoid_bf16r(gx.y));
                G[p] = __ffma2_rn(ls, make_float2(a.gate_c, a.gate_c), G[p]);
                const float2 eq = make_float2(ex2_approx(G[p].x + a.log2_scale), ex2_approx(G[p].y + a.log2_scale));
                const float2 ek = make_float2(ex2_approx(-G[p].x), ex2_approx(-G[p].y));
                const float2 qt = __fmul2_rn(qv, eq), kt = __fmul2_rn(kv, ek);
                oq[p] = pack_bf16(qt.x, qt.y);
                ok[p] = pack_bf16(kt.x, kt.y);
            }
            *reinterpret_cast<uint2 *>(qo + (size_t)l * a.Dk) = make_uint2(oq[0], oq[1]);
            *reinterpret_cast<uint2 *>(ko + (size_t)l * a.Dk) = make_uint2(ok[0], ok[1]);
        }
    }
    // (rows past the end of a partial last chunk contribute gk = 0, like the TMA zero fill of the in-kernel pre-pass)
    if (a.envelope_flag != nullptr &&
        fminf(fminf(G[0].x, G[0].y), fminf(G[1].x, G[1].y)) < GATE_ENVELOPE_LOG2)
        atomicOr(a.envelope_flag, 1);
    const int h = d0 / a.K, kap = d0 - h * a.K;
    const int H = a.Dk / a.K;
    *reinterpret_cast<float4 *>(a.decay + (((size_t)b * H + h) * a.NT + n) * a.K + kap) =
        make_float4(ex2_approx(G[0].x), ex2_approx(G[0].y), ex2_approx(G[1].x), ex2_approx(G[1].y));
    if (a.cq != nullptr && n == a.NT - 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int l = L - 4 + j;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float vq = l >= 0 ? __bfloat162float(xq[(size_t)l * a.ldq + c]) : 0.f;
                const float vk = l >= 0 ? __bfloat162float(xk[(size_t)l * a.ldk + c]) : 0.f;
                store_dyn(a.cq, a.cache_dtype, ((size_t)b * a.Dk + d0 + c) * 4 + j, vq);
                store_dyn(a.ck, a.cache_dtype, ((size_t)b * a.Dk + d0 + c) * 4 + j, vk);
            }
        }
    }
}

template <typename T>
int launch_prep(const PrepArgs &a, int nseg, cudaStream_t st) {
    constexpr int VEC = 16 / sizeof(T);
    const int TL = g_lina_variant[0] == 16 ? 16 : 8;
    const int tiles = (a.L + TL - 1) / TL;
    long long maxthreads = 0;
    for (int s = 0; s < nseg; ++s) {
        const long long n = (long long)a.B * tiles * (a.seg[s].D / VEC);
        if (n > maxthreads) maxthreads = n;
    }
    const long long nblk = (maxthreads + PREP_THREADS - 1) / PREP_THREADS;
    LINA_REQUIRE(nblk <= 2147483647LL, LINA_ERR_UNSUPPORTED, "gla_prep: grid too large");
    dim3 grid((unsigned)nblk, nseg);
    if (TL == 16) gla_prep_kernel<T, 16><<<grid, PREP_THREADS, 0, st>>>(a);
    else gla_prep_kernel<T, 8><<<grid, PREP_THREADS, 0, st>>>(a);
    LINA_LAUNCH_OK("gla_prep_kernel");
    return LINA_OK;
}

bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

}  // namespace

static int launch_conv4_bf16(const void *x, long long ldx, const void *w, void *y, void *cache, int cache_dtype, int B,
                             int L, int D, int silu, cudaStream_t st) {
    const int TL = g_lina_variant[0] == 16 ? 16 : 8;
    const long long nthreads = (long long)B * ((L + TL - 1) / TL) * (D / 8);
    const long long nblk = (nthreads + PREP_THREADS - 1) / PREP_THREADS;
    LINA_REQUIRE(nblk <= 2147483647LL, LINA_ERR_UNSUPPORTED, "short_conv: grid too large");
    if (TL == 16)
        conv4_silu_bf16_kernel<16><<<(unsigned)nblk, PREP_THREADS, 0, st>>>((const bf16 *)x, ldx, (const bf16 *)w, (bf16 *)y,
                                                                           cache, cache_dtype, B, L, D, silu);
    else
        conv4_silu_bf16_kernel<8><<<(unsigned)nblk, PREP_THREADS, 0, st>>>((const bf16 *)x, ldx, (const bf16 *)w, (bf16 *)y,
                                                                          cache, cache_dtype, B, L, D, silu);
    LINA_LAUNCH_OK("conv4_silu_bf16_kernel");
    return LINA_OK;
}

// single conv segment (used by lina_short_conv_fwd for W == 4 and 16-byte aligned rows)
int lina_short_conv4_tiles(const void *x, long long ldx, const void *w, void *y, void *cache, int cache_dtype, int B,
                           int L, int D, int silu, int dtype, void *stream) {
    if (dtype == LINA_BF16 && g_lina_variant[3] == 0)
        return launch_conv4_bf16(x, ldx, w, y, cache, cache_dtype, B, L, D, silu, (cudaStream_t)stream);
    PrepArgs a{};
    a.seg[0] = PrepSeg{x, w, y, cache, ldx, D, 0};
    a.B = B; a.L = L; a.cache_dtype = cache_dtype; a.silu = silu; a.use_clamp = 0; a.inv_norm = 1.f; a.clamp_min = 0.f;
    LINA_DISPATCH_DTYPE(dtype, return launch_prep<T_>(a, 1, (cudaStream_t)stream));
    return LINA_OK;
}

extern "C" int lina_gla_prefill_prep(const void *xq, const void *xk, const void *xv, long long ldx, const void *wq,
                                     const void *wk, const void *wv, const void *gk_raw, long long ldg, void *q, void *k,
                                     void *v, void *gk, void *cq, void *ck, void *cv, int cache_dtype, int B, int L,
                                     int Dk, int Dv, int W, float gate_normalizer, float clamp_min, int use_clamp,
                                     int dtype, void *stream) {
    LINA_REQUIRE(xq && xk && xv && wq && wk && wv && gk_raw && q && k && v && gk, LINA_ERR_BAD_ARG,
                 "gla_prefill_prep: null pointer");
    LINA_REQUIRE(B > 0 && L > 0 && Dk > 0 && Dv > 0 && gate_normalizer != 0.f, LINA_ERR_BAD_ARG,
                 "gla_prefill_prep: non-positive size");
    LINA_REQUIRE(lina_dtype_ok(dtype), LINA_ERR_BAD_ARG, "gla_prefill_prep: unknown dtype %d", dtype);
    LINA_REQUIRE(W == 4, LINA_ERR_UNSUPPORTED, "gla_prefill_prep: conv size %d (only 4, the shipped model's)", W);
    LINA_REQUIRE((cq == nullptr) == (ck == nullptr) && (cq == nullptr) == (cv == nullptr), LINA_ERR_BAD_ARG,
                 "gla_prefill_prep: pass all three conv caches or none");
    LINA_REQUIRE(cq == nullptr || lina_dtype_ok(cache_dtype), LINA_ERR_BAD_ARG, "gla_prefill_prep: bad cache dtype");
    const int vec = 16 / (int)lina_dtype_size(dtype);
    LINA_REQUIRE(Dk % vec == 0 && Dv % vec == 0 && ldx % vec == 0 && ldg % vec == 0 && ldx >= Dk && ldg >= Dk,
                 LINA_ERR_UNSUPPORTED, "gla_prefill_prep: channel counts / row strides must be multiples of %d", vec);
    LINA_REQUIRE(aligned16(xq) && aligned16(xk) && aligned16(xv) && aligned16(gk_raw) && aligned16(q) && aligned16(k) &&
                     aligned16(v) && aligned16(gk) && aligned16(wq) && aligned16(wk) && aligned16(wv),
                 LINA_ERR_UNSUPPORTED, "gla_prefill_prep: tensors must be 16-byte aligned");
    PrepArgs a{};
    a.seg[0] = PrepSeg{xv, wv, v, cv, ldx, Dv, 0};        // the widest segment first
    a.seg[1] = PrepSeg{xq, wq, q, cq, ldx, Dk, 0};
    a.seg[2] = PrepSeg{xk, wk, k, ck, ldx, Dk, 0};
    a.seg[3] = PrepSeg{gk_raw, nullptr, gk, nullptr, ldg, Dk, 1};
    a.B = B; a.L = L; a.cache_dtype = cache_dtype; a.silu = 1; a.use_clamp = use_clamp;
    a.inv_norm = 1.f / gate_normalizer; a.clamp_min = clamp_min;
    LINA_DISPATCH_DTYPE(dtype, return launch_prep<T_>(a, 4, (cudaStream_t)stream));
    return LINA_OK;
}

// q/k/v short convs with the chunk gating folded in (bf16, the tensor-core GLA kernel's operands): see qk_gate_bf16_kernel.
extern "C" int lina_gla_prefill_prep_gated(const void *xq, long long ldq, const void *xk, long long ldk, const void *xv,
                                           long long ldv, const void *wq,
                                           const void *wk, const void *wv, const void *gk_raw, long long ldg, void *qg,
                                           void *kg, void *v, float *decay, void *cq, void *ck, void *cv, int cache_dtype,
                                           int B, int L, int H, int K, int V, int W, float gate_normalizer, float scale,
                                           int *envelope_flag, void *stream) {
    LINA_REQUIRE(xq && xk && xv && wq && wk && wv && gk_raw && qg && kg && v && decay, LINA_ERR_BAD_ARG,
                 "gla_prefill_prep_gated: null pointer");
    LINA_REQUIRE(B > 0 && L > 0 && H > 0 && K > 0 && V > 0 && gate_normalizer != 0.f && scale > 0.f, LINA_ERR_BAD_ARG,
                 "gla_prefill_prep_gated: bad size / scale");
    LINA_REQUIRE(W == 4, LINA_ERR_UNSUPPORTED, "gla_prefill_prep_gated: conv size %d (only 4, the shipped model's)", W);
    LINA_REQUIRE((cq == nullptr) == (ck == nullptr) && (cq == nullptr) == (cv == nullptr), LINA_ERR_BAD_ARG,
                 "gla_prefill_prep_gated: pass all three conv caches or none");
    LINA_REQUIRE(cq == nullptr || lina_dtype_ok(cache_dtype), LINA_ERR_BAD_ARG, "gla_prefill_prep_gated: bad cache dtype");
    const int Dk = H * K, Dv = H * V;
    LINA_REQUIRE(K % 4 == 0 && Dv % 8 == 0 && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldg % 4 == 0 && ldq >= Dk &&
                     ldk >= Dk && ldv >= Dv && ldg >= Dk, LINA_ERR_UNSUPPORTED,
                 "gla_prefill_prep_gated: K %% 4, H*V %% 8 and 16-byte row strides required");
    LINA_REQUIRE(aligned16(xq) && aligned16(xk) && aligned16(xv) && aligned16(gk_raw) && aligned16(qg) && aligned16(kg) &&
                     aligned16(v) && aligned16(decay) && aligned16(wq) && aligned16(wk) && aligned16(wv),
                 LINA_ERR_UNSUPPORTED, "gla_prefill_prep_gated: tensors must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_conv4_bf16(xv, ldv, wv, v, cv, cache_dtype, B, L, Dv, 1, st);
    if (rc) return rc;
    GateArgs a{};
    a.xq = (const bf16 *)xq; a.xk = (const bf16 *)xk; a.graw = (const bf16 *)gk_raw;
    a.wq = (const bf16 *)wq; a.wk = (const bf16 *)wk; a.qg = (bf16 *)qg; a.kg = (bf16 *)kg; a.decay = decay;
    a.cq = cq; a.ck = ck; a.ldq = ldq; a.ldk = ldk; a.ldg = ldg; a.B = B; a.L = L; a.Dk = Dk; a.K = K; a.NT = (L + GC - 1) / GC;
    a.cache_dtype = cache_dtype;
    a.envelope_flag = envelope_flag;
    a.log2_scale = log2f(scale);
    a.gate_c = 1.44269504088896340736f / gate_normalizer;
    const long long nthreads = (long long)B * a.NT * (Dk / 4);
    const long long nblk = (nthreads + PREP_THREADS - 1) / PREP_THREADS;
    LINA_REQUIRE(nblk <= 2147483647LL, LINA_ERR_UNSUPPORTED, "gla_prefill_prep_gated: grid too large");
    qk_gate_bf16_kernel<<<(unsigned)nblk, PREP_THREADS, 0, st>>>(a);
    LINA_LAUNCH_OK("qk_gate_bf16_kernel");
    return LINA_OK;
}
