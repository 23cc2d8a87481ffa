// FusedRMSNormSwishGate: y = x * rsqrt(mean(x^2) + eps) * w * g * sigmoid(g) over rows of length N.
//
// Replaces FLA/fla/modules/fused_norm_gate.py:72-139 (fwd kernel) and :220-335 (bwd kernel), entry
// rms_norm_swish_gate_fn :439-518.  HBM streaming: fwd reads x,g and writes y once (3 tensors);
// bwd reads x,g,dy and writes dx,dg (5 tensors).  One warp per row, 16-byte vector accesses, the row
// stays in registers between the reduction and the normalisation (N <= 32*VPL*MAXCH elements).
#include "common.cuh"

namespace {

constexpr int MAXCH = 4;   // 16-byte chunks per lane kept in registers

template <typename T> struct V16 { static constexpr int n = 16 / sizeof(T); };

template <typename T> __device__ __forceinline__ void load16(const T *p, float *x) {
    constexpr int n = V16<T>::n;
    const uint4 raw = *reinterpret_cast<const uint4 *>(p);
    const T *e = reinterpret_cast<const T *>(&raw);
#pragma unroll
    for (int i = 0; i < n; ++i) x[i] = to_f(e[i]);
}
template <typename T> __device__ __forceinline__ void store16(T *p, const float *x) {
    constexpr int n = V16<T>::n;
    uint4 raw;
    T *e = reinterpret_cast<T *>(&raw);
#pragma unroll
    for (int i = 0; i < n; ++i) e[i] = from_f<T>(x[i]);
    *reinterpret_cast<uint4 *>(p) = raw;
}

// Forward.  Row r of x / y is contiguous; row r of g lives at g + (r / g_group) * ldg + (r % g_group) * N, so the gate can be
// read in place as a column slice of the [q;k;v;g] projection buffer (g_group = heads, ldg = its row stride) -- dense is
// g_group = 1, ldg = N.  Both the x and the g loads of a row are issued before the reduction (2 KB in flight per warp).
template <typename T, int CH>
__global__ void __launch_bounds__(256)
norm_gate_fwd_kernel(const T *__restrict__ x, const T *__restrict__ g, const T *__restrict__ w, T *__restrict__ y,
                     float *__restrict__ rstd_out, int M, int N, float eps, int g_group, long long ldg) {
    constexpr int n = V16<T>::n;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const T *xr = x + (size_t)row * N;
    const T *gr = g + (size_t)(row / g_group) * ldg + (size_t)(row % g_group) * N;
    T *yr = y + (size_t)row * N;
    const int nch = N / n;                 // host guarantees N % n == 0 and nch <= 32*CH
    uint4 xraw[CH], graw[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int ch = lane + c * 32;
        if (ch < nch) {
            xraw[c] = *reinterpret_cast<const uint4 *>(xr + (size_t)ch * n);
            graw[c] = *reinterpret_cast<const uint4 *>(gr + (size_t)ch * n);
        }
    }
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        if (lane + c * 32 < nch) {
            const T *e = reinterpret_cast<const T *>(&xraw[c]);
#pragma unroll
            for (int i = 0; i < n; ++i) { const float v = to_f(e[i]); ss = fmaf(v, v, ss); }
        }
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / (float)N + eps);
    if (rstd_out != nullptr && lane == 0) rstd_out[row] = rstd;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int ch = lane + c * 32;
        if (ch < nch) {
            float wv[n], out[n];
            if (w != nullptr) load16<T>(w + (size_t)ch * n, wv);
            const T *xe = reinterpret_cast<const T *>(&xraw[c]), *ge = reinterpret_cast<const T *>(&graw[c]);
#pragma unroll
            for (int i = 0; i < n; ++i) {
                const float gv = to_f(ge[i]);
                const float yh = to_f(xe[i]) * rstd * (w != nullptr ? wv[i] : 1.f);
                out[i] = yh * gv * sigmoid_io<T>(gv);
            }
            store16<T>(yr + (size_t)ch * n, out);
        }
    }
}

// dx = rstd * (dxh - xh * mean(dxh * xh)), dxh = dy * w * swish(g), xh = x * rstd
// dg = dy * xh * w * swish'(g), swish'(g) = s (1 + g (1 - s)) ;  dw += sum_rows dy * xh * swish(g)
// One warp per row at a time; the three row loads (x, g, dy) are issued before any use; dw is accumulated in registers over
// the warp's rows and flushed with one atomic per column.  CH = 16-byte chunks per lane (2 for the shipped V = 512 in
// bf16: 80 registers instead of 142, which had left one block per SM and 0.6 TB/s).
template <typename T, int CH>
__global__ void __launch_bounds__(256)
norm_gate_bwd_kernel(const T *__restrict__ x, const T *__restrict__ g, const T *__restrict__ w,
                     const float *__restrict__ rstd_in, const T *__restrict__ dy, T *__restrict__ dx,
                     T *__restrict__ dg, float *__restrict__ dw, int M, int N, int rows_per_warp) {
    constexpr int n = V16<T>::n;
    const int wid = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int nch = N / n;
    float dwl[CH][n], wv[CH][n];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int ch = lane + c * 32;
#pragma unroll
        for (int i = 0; i < n; ++i) { dwl[c][i] = 0.f; wv[c][i] = 1.f; }
        if (w != nullptr && ch < nch) load16<T>(w + (size_t)ch * n, wv[c]);
    }
    (void)rows_per_warp;
    for (int row = wid; row < M; row += gridDim.x * 8) {       // persistent: rows strided over all warps of the grid
        const size_t off = (size_t)row * N;
        const float rstd = rstd_in[row];
        uint4 xr[CH], gr[CH], dr[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int ch = lane + c * 32;
            if (ch < nch) {
                xr[c] = *reinterpret_cast<const uint4 *>(x + off + (size_t)ch * n);
                gr[c] = *reinterpret_cast<const uint4 *>(g + off + (size_t)ch * n);
                dr[c] = *reinterpret_cast<const uint4 *>(dy + off + (size_t)ch * n);
            }
        }
        float xh[CH][n], dxh[CH][n];
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int ch = lane + c * 32;
            if (ch < nch) {
                const T *xe = reinterpret_cast<const T *>(&xr[c]), *ge = reinterpret_cast<const T *>(&gr[c]);
                const T *de = reinterpret_cast<const T *>(&dr[c]);
                float dgo[n];
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    const float gv = to_f(ge[i]), dyv = to_f(de[i]);
                    const float s = sigmoidf_(gv);
                    const float sw = gv * s;
                    xh[c][i] = to_f(xe[i]) * rstd;
                    dxh[c][i] = dyv * wv[c][i] * sw;
                    dot = fmaf(dxh[c][i], xh[c][i], dot);
                    dgo[i] = dyv * xh[c][i] * wv[c][i] * s * (1.f + gv * (1.f - s));
                    dwl[c][i] = fmaf(dyv * xh[c][i], sw, dwl[c][i]);
                }
                store16<T>(dg + off + (size_t)ch * n, dgo);
            }
        }
        dot = warp_sum(dot) / (float)N;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int ch = lane + c * 32;
            if (ch < nch) {
                float out[n];
#pragma unroll
                for (int i = 0; i < n; ++i) out[i] = rstd * (dxh[c][i] - xh[c][i] * dot);
                store16<T>(dx + off + (size_t)ch * n, out);
            }
        }
    }
    if (dw != nullptr) {
        // block-level reduction over the 8 warps in shared memory, then ONE atomic per column per block (the grid is
        // 2 blocks per SM: ~150 k atomics in total; per-warp atomics onto N addresses serialised in L2 and cost more than
        // the streaming itself)
        __shared__ float red[8][32 * CH * n + 1];
        const int wl = threadIdx.x >> 5;
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
            for (int i = 0; i < n; ++i) red[wl][(lane + c * 32) * n + i] = dwl[c][i];
        __syncthreads();
        for (int col = threadIdx.x; col < N; col += 256) {
            float a = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) a += red[w8][col];
            atomicAdd(&dw[col], a);
        }
    }
}

int check(int M, int N, int dtype) {
    LINA_REQUIRE(M > 0 && N > 0, LINA_ERR_BAD_ARG, "rmsnorm_swishgate: non-positive size");
    LINA_REQUIRE(lina_dtype_ok(dtype), LINA_ERR_BAD_ARG, "rmsnorm_swishgate: unknown dtype %d", dtype);
    const int n = 16 / (int)lina_dtype_size(dtype);
    LINA_REQUIRE(N % n == 0 && N / n <= 32 * MAXCH, LINA_ERR_UNSUPPORTED,
                 "rmsnorm_swishgate: row length N=%d must be a multiple of %d and <= %d", N, n, 32 * MAXCH * n);
    return LINA_OK;
}

}  // namespace

extern "C" int lina_rmsnorm_swishgate_fwd_ld(const void *x, const void *g, const void *w, void *y, float *rstd, int M,
                                             int N, float eps, int g_group, long long ldg, int dtype, void *stream) {
    LINA_REQUIRE(x && g && y, LINA_ERR_BAD_ARG, "rmsnorm_swishgate_fwd: null pointer");
    int rc = check(M, N, dtype);
    if (rc) return rc;
    const int n = 16 / (int)lina_dtype_size(dtype);
    LINA_REQUIRE(g_group >= 1 && ldg >= (long long)g_group * N && ldg % n == 0 && M % g_group == 0, LINA_ERR_BAD_ARG,
                 "rmsnorm_swishgate_fwd: bad gate layout (g_group=%d, ldg=%lld, N=%d, M=%d)", g_group, ldg, N, M);
    LINA_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)g % 16 == 0 && (uintptr_t)y % 16 == 0 && (uintptr_t)w % 16 == 0,
                 LINA_ERR_UNSUPPORTED, "rmsnorm_swishgate_fwd: tensors must be 16-byte aligned");
    const int nch = N / n;
    if (nch <= 64) {
        LINA_DISPATCH_DTYPE(dtype, norm_gate_fwd_kernel<T_, 2><<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
                                       (const T_ *)x, (const T_ *)g, (const T_ *)w, (T_ *)y, rstd, M, N, eps, g_group, ldg));
    } else {
        LINA_DISPATCH_DTYPE(dtype, norm_gate_fwd_kernel<T_, MAXCH><<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
                                       (const T_ *)x, (const T_ *)g, (const T_ *)w, (T_ *)y, rstd, M, N, eps, g_group, ldg));
    }
    LINA_LAUNCH_OK("norm_gate_fwd_kernel");
    return LINA_OK;
}

extern "C" int lina_rmsnorm_swishgate_fwd(const void *x, const void *g, const void *w, void *y, float *rstd, int M,
                                          int N, float eps, int dtype, void *stream) {
    return lina_rmsnorm_swishgate_fwd_ld(x, g, w, y, rstd, M, N, eps, 1, N, dtype, stream);
}

extern "C" int lina_rmsnorm_swishgate_bwd(const void *x, const void *g, const void *w, const float *rstd,
                                          const void *dy, void *dx, void *dg, float *dw, int M, int N, int dtype,
                                          void *stream) {
    LINA_REQUIRE(x && g && rstd && dy && dx && dg, LINA_ERR_BAD_ARG, "rmsnorm_swishgate_bwd: null pointer");
    int rc = check(M, N, dtype);
    if (rc) return rc;
    // persistent grid: 2 blocks of 8 warps per SM, rows strided over the warps
    const int rows_per_warp = 0;
    int nwarps = (M + 7) / 8 * 8;
    if (nwarps > 148 * 2 * 8) nwarps = 148 * 2 * 8;
    const int nch = N / (16 / (int)lina_dtype_size(dtype));
    if (nch <= 64) {
        LINA_DISPATCH_DTYPE(dtype, norm_gate_bwd_kernel<T_, 2><<<(nwarps + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
                                       (const T_ *)x, (const T_ *)g, (const T_ *)w, rstd, (const T_ *)dy, (T_ *)dx,
                                       (T_ *)dg, dw, M, N, rows_per_warp));
    } else {
        LINA_DISPATCH_DTYPE(dtype, norm_gate_bwd_kernel<T_, MAXCH><<<(nwarps + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
                                       (const T_ *)x, (const T_ *)g, (const T_ *)w, rstd, (const T_ *)dy, (T_ *)dx,
                                       (T_ *)dg, dw, M, N, rows_per_warp));
    }
    LINA_LAUNCH_OK("norm_gate_bwd_kernel");
    return LINA_OK;
}
