// ShortConvolution: depthwise causal conv (kernel W <= 8) + SiLU over [B,L,D], channels-last.
//
// Replaces the external causal-conv1d 1.3.0.post1 CUDA kernels the reference calls at
// FLA/fla/modules/convolution.py:168-173 (causal_conv1d_fn) and :189-195 (causal_conv1d_update);
// semantics are those of the in-tree torch branch :175-178 / :197-204.
//
// Pure HBM streaming: x is read once (+ (W-1)/LT halo), y written once.  Threads run along the
// channel dim (coalesced rows of [B,L,D]); each thread slides a W-tap window over LT time steps.
#include "common.cuh"
#include "sm100.cuh"
#include "packed.cuh"

extern int g_lina_variant[16];
// gla_prep.cu: TL-row tiles, all loads of a thread issued up front (the production W == 4 path)
int lina_short_conv4_tiles(const void *x, long long ldx, const void *w, void *y, void *cache, int cache_dtype, int B,
                           int L, int D, int silu, int dtype, void *stream);

namespace {

constexpr int LT = 16;      // time steps per thread
constexpr int MAXW = 8;

template <typename T>
__global__ void __launch_bounds__(128)
short_conv_fwd_kernel(const T *__restrict__ x, const T *__restrict__ w, T *__restrict__ y, void *__restrict__ cache,
                      int cache_dtype, int L, int D, int W, int silu) {
    const int d = blockIdx.x * 128 + threadIdx.x;
    const int l0 = blockIdx.y * LT, b = blockIdx.z;
    if (d >= D) return;
    float wt[MAXW], win[MAXW];
#pragma unroll
    for (int j = 0; j < MAXW; ++j) wt[j] = j < W ? to_f(w[(size_t)d * W + j]) : 0.f;
    const T *xb = x + (size_t)b * L * D + d;
    T *yb = y + (size_t)b * L * D + d;
    // win[j] holds x[l - (W-1) + j]; preload the W-1 halo values
#pragma unroll
    for (int j = 0; j < MAXW; ++j) {
        const int l = l0 - (W - 1) + j;
        win[j] = (j < W - 1 && l >= 0) ? to_f(xb[(size_t)l * D]) : 0.f;
    }
    const int lend = min(L, l0 + LT);
    for (int l = l0; l < lend; ++l) {
        const float xv = to_f(xb[(size_t)l * D]);
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < MAXW; ++j) {
            if (j < W - 1) acc = fmaf(win[j], wt[j], acc);
        }
        // tap W-1 multiplies the newest sample
        float wl = 0.f;
#pragma unroll
        for (int j = 0; j < MAXW; ++j) if (j == W - 1) wl = wt[j];
        acc = fmaf(xv, wl, acc);
        yb[(size_t)l * D] = from_f<T>(silu ? siluf_(acc) : acc);
#pragma unroll
        for (int j = 0; j < MAXW - 1; ++j) {
            if (j < W - 2) win[j] = win[j + 1];
            else if (j == W - 2) win[j] = xv;
        }
    }
    // cache[b,d,:] = last W inputs, zero left-padded when L < W   (convolution.py:164-166)
    if (cache != nullptr && lend == L) {
        for (int j = 0; j < W; ++j) {
            const int l = L - W + j;
            store_dyn(cache, cache_dtype, ((size_t)b * D + d) * W + j, l >= 0 ? to_f(xb[(size_t)l * D]) : 0.f);
        }
    }
}

// Vectorised variant for the shipped configuration (W = 4, D a multiple of the 16-byte vector): each thread
// owns 16 bytes of channels and slides over LTV time steps -> 512 B contiguous per warp per row.
constexpr int LTV = 32;
template <typename T>
__global__ void __launch_bounds__(128)
short_conv_fwd_w4_kernel(const T *__restrict__ x, const T *__restrict__ w, T *__restrict__ y, void *__restrict__ cache,
                         int cache_dtype, int L, int D, int silu) {
    constexpr int VEC = 16 / sizeof(T);
    const int dv = blockIdx.x * 128 + threadIdx.x;          // vector index along D
    const int d0 = dv * VEC;
    const int l0 = blockIdx.y * LTV, b = blockIdx.z;
    if (d0 >= D) return;
    float wt[4][VEC], win[3][VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) wt[j][c] = to_f(w[(size_t)(d0 + c) * 4 + j]);
    const T *xb = x + (size_t)b * L * D + d0;
    T *yb = y + (size_t)b * L * D + d0;
    auto load = [&](int l, float *out) {
        if (l < 0) {
#pragma unroll
            for (int c = 0; c < VEC; ++c) out[c] = 0.f;
            return;
        }
        const uint4 raw = *reinterpret_cast<const uint4 *>(xb + (size_t)l * D);
        const T *e = reinterpret_cast<const T *>(&raw);
#pragma unroll
        for (int c = 0; c < VEC; ++c) out[c] = to_f(e[c]);
    };
    load(l0 - 3, win[0]); load(l0 - 2, win[1]); load(l0 - 1, win[2]);
    const int lend = min(L, l0 + LTV);
#pragma unroll 4
    for (int l = l0; l < lend; ++l) {
        float xv[VEC];
        load(l, xv);
        uint4 raw;
        T *e = reinterpret_cast<T *>(&raw);
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
            float acc = win[0][c] * wt[0][c];
            acc = fmaf(win[1][c], wt[1][c], acc);
            acc = fmaf(win[2][c], wt[2][c], acc);
            acc = fmaf(xv[c], wt[3][c], acc);
            e[c] = from_f<T>(silu ? acc * sigmoid_io<T>(acc) : acc);
            win[0][c] = win[1][c]; win[1][c] = win[2][c]; win[2][c] = xv[c];
        }
        *reinterpret_cast<uint4 *>(yb + (size_t)l * D) = raw;
    }
    if (cache != nullptr && lend == L) {
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
            const size_t cb = ((size_t)b * D + d0 + c) * 4;
            if (L >= 4) {
                store_dyn(cache, cache_dtype, cb + 0, to_f(xb[(size_t)(L - 4) * D + c]));
            } else {
                store_dyn(cache, cache_dtype, cb + 0, 0.f);
            }
            store_dyn(cache, cache_dtype, cb + 1, L >= 3 ? win[0][c] : 0.f);
            store_dyn(cache, cache_dtype, cb + 2, L >= 2 ? win[1][c] : 0.f);
            store_dyn(cache, cache_dtype, cb + 3, win[2][c]);
        }
    }
}

// dx[l] = sum_j w[j] * dpre[l + (W-1) - j],  dpre = dy * act'(pre) ; dw[j] += sum_l dpre[l] * x[l-(W-1)+j]
template <typename T>
__global__ void __launch_bounds__(128)
short_conv_bwd_kernel(const T *__restrict__ x, const T *__restrict__ w, const T *__restrict__ dy,
                      T *__restrict__ dx, float *__restrict__ dw, int L, int D, int W, int silu) {
    const int d = blockIdx.x * 128 + threadIdx.x;
    const int l0 = blockIdx.y * LT, b = blockIdx.z;
    if (d >= D) return;
    float wt[MAXW], dwl[MAXW];
#pragma unroll
    for (int j = 0; j < MAXW; ++j) { wt[j] = j < W ? to_f(w[(size_t)d * W + j]) : 0.f; dwl[j] = 0.f; }
    const T *xb = x + (size_t)b * L * D + d;
    const T *dyb = dy + (size_t)b * L * D + d;
    T *dxb = dx + (size_t)b * L * D + d;
    auto xat = [&](int l) -> float { return (l >= 0 && l < L) ? to_f(xb[(size_t)l * D]) : 0.f; };
    auto dpre_at = [&](int l) -> float {
        if (l < 0 || l >= L) return 0.f;
        const float g = to_f(dyb[(size_t)l * D]);
        if (!silu) return g;
        float pre = 0.f;
        for (int j = 0; j < W; ++j) pre = fmaf(xat(l - (W - 1) + j), wt[j], pre);
        const float s = sigmoidf_(pre);
        return g * s * (1.f + pre * (1.f - s));
    };
    const int lend = min(L, l0 + LT);
    for (int l = l0; l < lend; ++l) {
        float acc = 0.f;
        for (int j = 0; j < W; ++j) acc = fmaf(wt[j], dpre_at(l + (W - 1) - j), acc);
        dxb[(size_t)l * D] = from_f<T>(acc);
        const float dp = dpre_at(l);
        for (int j = 0; j < W; ++j) dwl[j] = fmaf(dp, xat(l - (W - 1) + j), dwl[j]);
    }
    for (int j = 0; j < W; ++j) atomicAdd(&dw[(size_t)d * W + j], dwl[j]);
}


// Backward for the shipped configuration (bf16, W = 4), packed math, one pass:
//   pre[m] = sum_j w[j] x[m-3+j] ;  dpre[m] = dy[m] * act'(pre[m]) ;  dx[l] = sum_j w[j] dpre[l+3-j] ;  dw[j] += dpre[m] x[m-3+j]
// One thread = 4 channels x BRT consecutive rows, walked once with a 3-row x window and a 3-row dpre window (dx[m-3] is
// emitted when dpre[m] is known); rows l0+BRT .. l0+BRT+2 are recomputed as a tail halo.  dw is reduced in registers
// over the thread's rows and then with one fp32 atomic per (channel, tap).  The round-1 kernel recomputed pre four times
// per output with scalar loads (1.7 ms per call at bs8 x seq4096: 22 % of a training step).
constexpr int BRT = 64;      // rows per thread
constexpr int BRB = 4;       // rows per load batch
__global__ void __launch_bounds__(128, 4)
short_conv4_bwd_bf16_kernel(const bf16 *__restrict__ x, const bf16 *__restrict__ w, const bf16 *__restrict__ dy,
                            bf16 *__restrict__ dx, float *__restrict__ dw, int B, int L, int D, int silu) {
    const int ng = D / 4;
    const int tiles = (L + BRT - 1) / BRT;
    const long long idx = (long long)blockIdx.x * 128 + threadIdx.x;
    const int cg = (int)(idx % ng);
    const long long rest = idx / ng;
    const int tile = (int)(rest % tiles);
    const long long b = rest / tiles;
    if (b >= B) return;
    const int d0 = cg * 4, l0 = tile * BRT;
    const bf16 *xb = x + (size_t)b * L * D + d0, *dyb = dy + (size_t)b * L * D + d0;
    bf16 *dxb = dx + (size_t)b * L * D + d0;
    float2 wt[2][4];
    load_taps2(w + (size_t)d0 * 4, wt[0]);
    load_taps2(w + (size_t)(d0 + 2) * 4, wt[1]);
    float2 xw[2][3], dp[2][3], dwa[2][4];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dwa[p][j] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 3; ++j) dp[p][j] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int l = l0 - 3 + i;
        const uint2 r = l >= 0 ? *reinterpret_cast<const uint2 *>(xb + (size_t)l * D) : make_uint2(0, 0);
        xw[0][i] = bf2_to_f2(r.x); xw[1][i] = bf2_to_f2(r.y);
    }
    const int mend = min(L + 3, l0 + BRT + 3);            // exclusive; rows >= L contribute dpre = 0
#pragma unroll 1
    for (int m0 = l0; m0 < mend; m0 += BRB) {
        uint2 rx[BRB], rd[BRB];
#pragma unroll
        for (int i = 0; i < BRB; ++i) {
            const int m = m0 + i;
            const bool ok = m < L && m < mend;
            rx[i] = ok ? *reinterpret_cast<const uint2 *>(xb + (size_t)m * D) : make_uint2(0, 0);
            rd[i] = ok ? *reinterpret_cast<const uint2 *>(dyb + (size_t)m * D) : make_uint2(0, 0);
        }
#pragma unroll
        for (int i = 0; i < BRB; ++i) {
            const int m = m0 + i;
            if (m >= mend) break;
            const bool own = m < l0 + BRT;                // rows of this tile (dw counted once); later rows are halo
            uint32_t ow[2];
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const float2 xv = bf2_to_f2(p == 0 ? rx[i].x : rx[i].y);
                const float2 g = bf2_to_f2(p == 0 ? rd[i].x : rd[i].y);
                // pre[m] from the window (x[m-3], x[m-2], x[m-1]) and x[m]
                float2 pre = __fmul2_rn(xw[p][0], wt[p][0]);
                pre = __ffma2_rn(xw[p][1], wt[p][1], pre);
                pre = __ffma2_rn(xw[p][2], wt[p][2], pre);
                pre = __ffma2_rn(xv, wt[p][3], pre);
                float2 d = g;
                if (silu) {                               // act'(pre) = s (1 + pre (1 - s)), s = sigmoid(pre)
                    const float2 hh = __fmul2_rn(pre, make_float2(0.5f, 0.5f));
                    const float2 t = make_float2(tanh_approx_(hh.x), tanh_approx_(hh.y));
                    const float2 sg = __ffma2_rn(t, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
                    const float2 om = __ffma2_rn(sg, make_float2(-1.f, -1.f), make_float2(1.f, 1.f));
                    const float2 u = __ffma2_rn(pre, om, make_float2(1.f, 1.f));
                    d = __fmul2_rn(g, __fmul2_rn(sg, u));
                }
                if (own) {                                // dw[j] += dpre[m] * x[m-3+j]
                    dwa[p][0] = __ffma2_rn(d, xw[p][0], dwa[p][0]);
                    dwa[p][1] = __ffma2_rn(d, xw[p][1], dwa[p][1]);
                    dwa[p][2] = __ffma2_rn(d, xw[p][2], dwa[p][2]);
                    dwa[p][3] = __ffma2_rn(d, xv, dwa[p][3]);
                }
                // dx[m-3] = w0 dpre[m] + w1 dpre[m-1] + w2 dpre[m-2] + w3 dpre[m-3]
                float2 o = __fmul2_rn(dp[p][0], wt[p][3]);
                o = __ffma2_rn(dp[p][1], wt[p][2], o);
                o = __ffma2_rn(dp[p][2], wt[p][1], o);
                o = __ffma2_rn(d, wt[p][0], o);
                ow[p] = sm100::pack_bf16(o.x, o.y);
                xw[p][0] = xw[p][1]; xw[p][1] = xw[p][2]; xw[p][2] = xv;
                dp[p][0] = dp[p][1]; dp[p][1] = dp[p][2]; dp[p][2] = d;
            }
            const int l = m - 3;
            if (l >= l0 && l < L) *reinterpret_cast<uint2 *>(dxb + (size_t)l * D) = make_uint2(ow[0], ow[1]);
        }
    }
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&dw[(size_t)(d0 + 2 * p) * 4 + j], dwa[p][j].x);
            atomicAdd(&dw[(size_t)(d0 + 2 * p + 1) * 4 + j], dwa[p][j].y);
        }
}

template <typename T>
__global__ void short_conv_update_kernel(const T *__restrict__ x, void *__restrict__ cache, int cache_dtype,
                                         const T *__restrict__ w, T *__restrict__ y, int B, int D, int W,
                                         int silu) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * D) return;
    const int d = (int)(i % D);
    const size_t cbase = (size_t)i * W;
    float acc = 0.f;
    for (int j = 0; j < W - 1; ++j) {
        const float nxt = load_dyn(cache, cache_dtype, cbase + j + 1);
        store_dyn(cache, cache_dtype, cbase + j, nxt);
        acc = fmaf(nxt, to_f(w[(size_t)d * W + j]), acc);
    }
    const float xv = to_f(x[i]);
    store_dyn(cache, cache_dtype, cbase + W - 1, xv);
    const float last = load_dyn(cache, cache_dtype, cbase + W - 1);
    acc = fmaf(last, to_f(w[(size_t)d * W + W - 1]), acc);
    y[i] = from_f<T>(silu ? siluf_(acc) : acc);
}

}  // namespace

extern "C" int lina_short_conv_fwd(const void *x, const void *w, void *y, void *cache, int cache_dtype, int B,
                                   int L, int D, int W, int silu, int dtype, void *stream) {
    LINA_REQUIRE(x && w && y, LINA_ERR_BAD_ARG, "short_conv_fwd: null pointer");
    LINA_REQUIRE(B > 0 && L > 0 && D > 0, LINA_ERR_BAD_ARG, "short_conv_fwd: non-positive size");
    LINA_REQUIRE(W >= 1 && W <= MAXW, LINA_ERR_UNSUPPORTED, "short_conv_fwd: kernel size %d not in [1,%d]", W, MAXW);
    LINA_REQUIRE(cache == nullptr || lina_dtype_ok(cache_dtype), LINA_ERR_BAD_ARG, "short_conv_fwd: bad cache dtype");
    LINA_REQUIRE(B <= 65535 && (L + LT - 1) / LT <= 65535, LINA_ERR_UNSUPPORTED, "short_conv_fwd: grid too large");
    const int vec = 16 / (int)lina_dtype_size(dtype);
    if (W == 4 && D % vec == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)w % 16 == 0) &&
        g_lina_variant[1] == 0)
        return lina_short_conv4_tiles(x, D, w, y, cache, cache_dtype, B, L, D, silu, dtype, stream);
    if (W == 4 && D % vec == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0)) {   // A/B: the round-1 sliding-window kernel
        const int nv = D / vec;
        dim3 gridv((nv + 127) / 128, (L + LTV - 1) / LTV, B);
        LINA_DISPATCH_DTYPE(dtype, short_conv_fwd_w4_kernel<T_><<<gridv, 128, 0, (cudaStream_t)stream>>>(
                                       (const T_ *)x, (const T_ *)w, (T_ *)y, cache, cache_dtype, L, D, silu));
        LINA_LAUNCH_OK("short_conv_fwd_w4_kernel");
        return LINA_OK;
    }
    dim3 grid((D + 127) / 128, (L + LT - 1) / LT, B);
    LINA_DISPATCH_DTYPE(dtype, short_conv_fwd_kernel<T_><<<grid, 128, 0, (cudaStream_t)stream>>>(
                                   (const T_ *)x, (const T_ *)w, (T_ *)y, cache, cache_dtype, L, D, W, silu));
    LINA_LAUNCH_OK("short_conv_fwd_kernel");
    return LINA_OK;
}

extern "C" int lina_short_conv_bwd(const void *x, const void *w, const void *dy, void *dx, float *dw, int B, int L,
                                   int D, int W, int silu, int dtype, void *stream) {
    LINA_REQUIRE(x && w && dy && dx && dw, LINA_ERR_BAD_ARG, "short_conv_bwd: null pointer");
    LINA_REQUIRE(B > 0 && L > 0 && D > 0, LINA_ERR_BAD_ARG, "short_conv_bwd: non-positive size");
    LINA_REQUIRE(W >= 1 && W <= MAXW, LINA_ERR_UNSUPPORTED, "short_conv_bwd: kernel size %d not in [1,%d]", W, MAXW);
    if (dtype == LINA_BF16 && W == 4 && D % 4 == 0 && ((uintptr_t)x % 8 == 0) && ((uintptr_t)dy % 8 == 0) &&
        ((uintptr_t)dx % 8 == 0) && ((uintptr_t)w % 16 == 0) && g_lina_variant[5] == 0) {
        const long long nthreads = (long long)B * ((L + BRT - 1) / BRT) * (D / 4);
        const long long nblk = (nthreads + 127) / 128;
        LINA_REQUIRE(nblk <= 2147483647LL, LINA_ERR_UNSUPPORTED, "short_conv_bwd: grid too large");
        short_conv4_bwd_bf16_kernel<<<(unsigned)nblk, 128, 0, (cudaStream_t)stream>>>(
            (const bf16 *)x, (const bf16 *)w, (const bf16 *)dy, (bf16 *)dx, dw, B, L, D, silu);
        LINA_LAUNCH_OK("short_conv4_bwd_bf16_kernel");
        return LINA_OK;
    }
    LINA_REQUIRE(B <= 65535 && (L + LT - 1) / LT <= 65535, LINA_ERR_UNSUPPORTED, "short_conv_bwd: grid too large");
    dim3 grid((D + 127) / 128, (L + LT - 1) / LT, B);
    LINA_DISPATCH_DTYPE(dtype, short_conv_bwd_kernel<T_><<<grid, 128, 0, (cudaStream_t)stream>>>(
                                   (const T_ *)x, (const T_ *)w, (const T_ *)dy, (T_ *)dx, dw, L, D, W, silu));
    LINA_LAUNCH_OK("short_conv_bwd_kernel");
    return LINA_OK;
}

extern "C" int lina_short_conv_update(const void *x, void *cache, int cache_dtype, const void *w, void *y, int B,
                                      int D, int W, int silu, int dtype, void *stream) {
    LINA_REQUIRE(x && cache && w && y, LINA_ERR_BAD_ARG, "short_conv_update: null pointer");
    LINA_REQUIRE(B > 0 && D > 0 && W >= 1 && W <= 64, LINA_ERR_BAD_ARG, "short_conv_update: bad size");
    LINA_REQUIRE(lina_dtype_ok(cache_dtype), LINA_ERR_BAD_ARG, "short_conv_update: bad cache dtype");
    const long long n = (long long)B * D;
    LINA_DISPATCH_DTYPE(dtype, short_conv_update_kernel<T_><<<(unsigned)((n + 255) / 256), 256, 0,
                                                              (cudaStream_t)stream>>>(
                                   (const T_ *)x, cache, cache_dtype, (const T_ *)w, (T_ *)y, B, D, W, silu));
    LINA_LAUNCH_OK("short_conv_update_kernel");
    return LINA_OK;
}
