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

int g_lina_variant[8] = {0, 0, 0, 0, 0, 0, 0, 0};

extern "C" int lina_debug_set_variant(int key, int value) {
    if (key < 0 || key >= 8) return LINA_ERR_BAD_ARG;
    g_lina_variant[key] = value;
    return LINA_OK;
}

namespace {

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

// single conv segment (used by lina_short_conv_fwd for W == 4 and 16-byte aligned rows)
int lina_short_conv4_tiles(const void *x, long long ldx, const void *w, void *y, void *cache, int cache_dtype, int B,
                           int L, int D, int silu, int dtype, void *stream) {
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
