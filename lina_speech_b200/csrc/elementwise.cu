// Small fused element-wise passes around the GLA mixer (HBM streaming, 16-byte vectors).
//   lina_gate_logsigmoid : gk = logsigmoid(x) / normalizer [clamped]           (model/gla.py:174-181)
//   lina_swiglu_act      : out[m, j] = silu(h[m, j]) * h[m, Hp + j]             (model/base_blocks.py:48-50)
//   lina_add_layernorm   : s = a + x ; y = LayerNorm(s)   (the residual add + pre-LN pairs of MixingBlock,
//                          model/base_blocks.py:65-68)
#include "common.cuh"

namespace {

// log(sigmoid(x)) with fast intrinsics; the result is rounded to the activation dtype right after, so the
// ~1e-7 absolute error of __logf/__expf is invisible
__device__ __forceinline__ float logsigmoid_fast(float x) { return fminf(x, 0.f) - __logf(1.f + __expf(-fabsf(x))); }

template <typename T>
__global__ void __launch_bounds__(256)
gate_logsigmoid_kernel(const T *__restrict__ x, T *__restrict__ y, long long n, float inv_norm, float clamp_min,
                       int use_clamp) {
    constexpr int VEC = 16 / sizeof(T);
    const long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * VEC;
    if (i >= n) return;
    if (i + VEC <= n) {
        const uint4 raw = *reinterpret_cast<const uint4 *>(x + i);
        const T *e = reinterpret_cast<const T *>(&raw);
        uint4 outr;
        T *oe = reinterpret_cast<T *>(&outr);
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
            // the reference rounds logsigmoid to the activation dtype before the division (a power of two)
            float g = to_f(from_f<T>(sizeof(T) == 4 ? logsigmoidf_(to_f(e[c])) : logsigmoid_fast(to_f(e[c])))) * inv_norm;
            if (use_clamp) g = fmaxf(g, clamp_min);
            oe[c] = from_f<T>(g);
        }
        *reinterpret_cast<uint4 *>(y + i) = outr;
    } else {
        for (long long j = i; j < n; ++j) {
            float g = to_f(from_f<T>(logsigmoidf_(to_f(x[j])))) * inv_norm;
            if (use_clamp) g = fmaxf(g, clamp_min);
            y[j] = from_f<T>(g);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
swiglu_act_kernel(const T *__restrict__ h, T *__restrict__ out, int M, int Hp) {
    constexpr int VEC = 16 / sizeof(T);
    const int nv = Hp / VEC;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)M * nv) return;
    const int m = (int)(idx / nv), j = (int)(idx - (long long)m * nv) * VEC;
    const uint4 graw = *reinterpret_cast<const uint4 *>(h + (size_t)m * 2 * Hp + j);
    const uint4 uraw = *reinterpret_cast<const uint4 *>(h + (size_t)m * 2 * Hp + Hp + j);
    const T *ge = reinterpret_cast<const T *>(&graw), *ue = reinterpret_cast<const T *>(&uraw);
    uint4 outr;
    T *oe = reinterpret_cast<T *>(&outr);
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
        const float g = to_f(ge[c]);
        oe[c] = from_f<T>(to_f(from_f<T>(siluf_(g))) * to_f(ue[c]));
    }
    *reinterpret_cast<uint4 *>(out + (size_t)m * Hp + j) = outr;
}

// sum = a + x (rounded to T like the reference's separate add), ln = LayerNorm(sum) * gamma + beta.  One warp per row,
// the row stays in registers between the statistics and the normalisation: 2 reads + 2 writes of the row in total
// (torch: add = 2r+1w, layer_norm = 1r+1w, in two launches).  a == nullptr: plain LayerNorm of x.
constexpr int LN_MAXCH = 8;
template <typename T, int LN_CH>
__global__ void __launch_bounds__(256)
add_layernorm_kernel(const T *__restrict__ a, const T *__restrict__ x, const T *__restrict__ gamma,
                     const T *__restrict__ beta, T *__restrict__ sum_out, T *__restrict__ ln_out, int M, int N, float eps) {
    constexpr int VEC = 16 / sizeof(T);
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const int nch = N / VEC;
    const size_t off = (size_t)row * N;
    float v[LN_CH][VEC];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < LN_CH; ++c) {
        const int ch = lane + c * 32;
        if (ch < nch) {
            const uint4 xr = *reinterpret_cast<const uint4 *>(x + off + (size_t)ch * VEC);
            const T *xe = reinterpret_cast<const T *>(&xr);
            if (a != nullptr) {
                const uint4 ar = *reinterpret_cast<const uint4 *>(a + off + (size_t)ch * VEC);
                const T *ae = reinterpret_cast<const T *>(&ar);
                uint4 sr;
                T *se = reinterpret_cast<T *>(&sr);
#pragma unroll
                for (int i = 0; i < VEC; ++i) { se[i] = from_f<T>(to_f(ae[i]) + to_f(xe[i])); v[c][i] = to_f(se[i]); }
                *reinterpret_cast<uint4 *>(sum_out + off + (size_t)ch * VEC) = sr;
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) v[c][i] = to_f(xe[i]);
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) s += v[c][i];
        }
    }
    // the affine parameters are fetched BEFORE the two dependent reductions: at decode-step sizes (a few dozen rows) the kernel is
    // one memory round trip long, and loading them after the statistics made it two
    uint4 gr[LN_CH], br[LN_CH];
#pragma unroll
    for (int c = 0; c < LN_CH; ++c) {
        const int ch = lane + c * 32;
        if (ch < nch) {
            gr[c] = *reinterpret_cast<const uint4 *>(gamma + (size_t)ch * VEC);
            br[c] = *reinterpret_cast<const uint4 *>(beta + (size_t)ch * VEC);
        }
    }
    const float mean = warp_sum(s) / (float)N;
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < LN_CH; ++c) {
        if (lane + c * 32 < nch) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) { const float d = v[c][i] - mean; ss = fmaf(d, d, ss); }
        }
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)N + eps);
#pragma unroll
    for (int c = 0; c < LN_CH; ++c) {
        const int ch = lane + c * 32;
        if (ch < nch) {
            const T *ge = reinterpret_cast<const T *>(&gr[c]), *be = reinterpret_cast<const T *>(&br[c]);
            uint4 outr;
            T *oe = reinterpret_cast<T *>(&outr);
#pragma unroll
            for (int i = 0; i < VEC; ++i) oe[i] = from_f<T>((v[c][i] - mean) * rstd * to_f(ge[i]) + to_f(be[i]));
            *reinterpret_cast<uint4 *>(ln_out + off + (size_t)ch * VEC) = outr;
        }
    }
}


// Row-wise cross entropy on logits of the activation dtype, fp32 math (the reference casts to float first:
// model/modeling_lina.py:104-106 F.cross_entropy(logits.float(), target, ignore_index=1)):
//   loss[m] = logsumexp(logits[m, :Vn]) - logits[m, target[m]]   (0 and valid[m] = 0 when target == ignore or masked out)
// One warp per row, single pass with an online (max, sum) pair per lane; rows may be strided (ld elements), e.g. the
// first Vn columns of a vocabulary-padded GEMM output.
template <typename T>
__global__ void __launch_bounds__(256)
cross_entropy_rows_kernel(const T *__restrict__ logits, long long ld, const long long *__restrict__ target,
                          const unsigned char *__restrict__ row_mask, float *__restrict__ loss, float *__restrict__ valid,
                          int M, int Vn, long long ignore_index) {
    constexpr int VEC = 16 / sizeof(T);
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const long long tg = target[row];
    const bool on = tg != ignore_index && (row_mask == nullptr || row_mask[row] != 0);
    if (!on || tg < 0 || tg >= Vn) {        // whole warp takes the same branch
        if (lane == 0) { loss[row] = 0.f; valid[row] = 0.f; }
        return;
    }
    const T *r = logits + (size_t)row * ld;
    const bool vec_ok = (ld % VEC == 0) && (((uintptr_t)logits & 15u) == 0);
    constexpr float LOG2E = 1.44269504088896340736f;
    float m = -INFINITY, s = 0.f;           // running max (natural units) and sum of exp(x - m)
    const int nfull = vec_ok ? Vn / VEC : 0;
    for (int ch0 = 0; ch0 < nfull; ch0 += 128) {          // 4 independent 16-byte loads per lane per iteration
        uint4 raw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int ch = ch0 + u * 32 + lane;
            if (ch < nfull) raw[u] = *reinterpret_cast<const uint4 *>(r + (size_t)ch * VEC);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int ch = ch0 + u * 32 + lane;
            if (ch < nfull) {
                const T *e = reinterpret_cast<const T *>(&raw[u]);
                float f[VEC], mx = -INFINITY;
#pragma unroll
                for (int i = 0; i < VEC; ++i) { f[i] = to_f(e[i]); mx = fmaxf(mx, f[i]); }
                if (mx > m) { s *= exp2f((m - mx) * LOG2E); m = mx; }
#pragma unroll
                for (int i = 0; i < VEC; ++i) s += exp2f((f[i] - m) * LOG2E);
            }
        }
    }
    for (int j = nfull * VEC + lane; j < Vn; j += 32) {   // tail (and the whole row when unaligned)
        const float f = to_f(r[j]);
        if (f > m) { s *= exp2f((m - f) * LOG2E); m = f; }
        s += exp2f((f - m) * LOG2E);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        const float mn = fmaxf(m, m2);
        const float a = (m == -INFINITY) ? 0.f : s * exp2f((m - mn) * LOG2E);
        const float b = (m2 == -INFINITY) ? 0.f : s2 * exp2f((m2 - mn) * LOG2E);
        s = a + b; m = mn;
    }
    if (lane == 0) {
        loss[row] = m + logf(s) - to_f(r[tg]);
        valid[row] = 1.f;
    }
}

}  // namespace

extern "C" int lina_add_layernorm(const void *a, const void *x, const void *gamma, const void *beta, void *sum_out,
                                  void *ln_out, int M, int N, float eps, int dtype, void *stream) {
    LINA_REQUIRE(x && gamma && beta && ln_out && M > 0 && N > 0, LINA_ERR_BAD_ARG, "add_layernorm: bad argument");
    LINA_REQUIRE(a == nullptr || sum_out != nullptr, LINA_ERR_BAD_ARG, "add_layernorm: sum_out required with a");
    LINA_REQUIRE(lina_dtype_ok(dtype), LINA_ERR_BAD_ARG, "add_layernorm: unknown dtype");
    const int vec = 16 / (int)lina_dtype_size(dtype);
    LINA_REQUIRE(N % vec == 0 && N / vec <= 32 * LN_MAXCH, LINA_ERR_UNSUPPORTED,
                 "add_layernorm: row length N=%d must be a multiple of %d and <= %d", N, vec, 32 * LN_MAXCH * vec);
    if (N / vec <= 32 * 4) {          // d_model 1024 in bf16: half the registers of the general case
        LINA_DISPATCH_DTYPE(dtype, add_layernorm_kernel<T_, 4><<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
                                       (const T_ *)a, (const T_ *)x, (const T_ *)gamma, (const T_ *)beta, (T_ *)sum_out,
                                       (T_ *)ln_out, M, N, eps));
    } else {
        LINA_DISPATCH_DTYPE(dtype, add_layernorm_kernel<T_, LN_MAXCH><<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
                                       (const T_ *)a, (const T_ *)x, (const T_ *)gamma, (const T_ *)beta, (T_ *)sum_out,
                                       (T_ *)ln_out, M, N, eps));
    }
    LINA_LAUNCH_OK("add_layernorm_kernel");
    return LINA_OK;
}

extern "C" int lina_gate_logsigmoid(const void *x, void *y, long long n, float normalizer, float clamp_min,
                                    int use_clamp, int dtype, void *stream) {
    LINA_REQUIRE(x && y && n > 0 && normalizer != 0.f, LINA_ERR_BAD_ARG, "gate_logsigmoid: bad argument");
    LINA_REQUIRE(lina_dtype_ok(dtype), LINA_ERR_BAD_ARG, "gate_logsigmoid: unknown dtype");
    LINA_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0, LINA_ERR_BAD_ARG, "gate_logsigmoid: 16-byte alignment");
    const int vec = 16 / (int)lina_dtype_size(dtype);
    const long long nthreads = (n + vec - 1) / vec;
    LINA_DISPATCH_DTYPE(dtype, gate_logsigmoid_kernel<T_><<<(unsigned)((nthreads + 255) / 256), 256, 0,
                                                             (cudaStream_t)stream>>>(
                                   (const T_ *)x, (T_ *)y, n, 1.f / normalizer, clamp_min, use_clamp));
    LINA_LAUNCH_OK("gate_logsigmoid_kernel");
    return LINA_OK;
}

extern "C" int lina_swiglu_act(const void *h, void *out, int M, int Hp, int dtype, void *stream) {
    LINA_REQUIRE(h && out && M > 0 && Hp > 0, LINA_ERR_BAD_ARG, "swiglu_act: bad argument");
    LINA_REQUIRE(lina_dtype_ok(dtype), LINA_ERR_BAD_ARG, "swiglu_act: unknown dtype");
    const int vec = 16 / (int)lina_dtype_size(dtype);
    LINA_REQUIRE(Hp % vec == 0 && (uintptr_t)h % 16 == 0 && (uintptr_t)out % 16 == 0, LINA_ERR_UNSUPPORTED,
                 "swiglu_act: hidden size %d must be a multiple of %d and pointers 16-byte aligned", Hp, vec);
    const long long nthreads = (long long)M * (Hp / vec);
    LINA_DISPATCH_DTYPE(dtype, swiglu_act_kernel<T_><<<(unsigned)((nthreads + 255) / 256), 256, 0,
                                                        (cudaStream_t)stream>>>((const T_ *)h, (T_ *)out, M, Hp));
    LINA_LAUNCH_OK("swiglu_act_kernel");
    return LINA_OK;
}

extern "C" int lina_cross_entropy_rows(const void *logits, long long ld, const int64_t *target, const uint8_t *row_mask,
                                       float *loss, float *valid, int M, int Vn, long long ignore_index, int dtype,
                                       void *stream) {
    LINA_REQUIRE(logits && target && loss && valid && M > 0 && Vn > 0 && ld >= Vn, LINA_ERR_BAD_ARG,
                 "cross_entropy_rows: bad argument");
    LINA_REQUIRE(lina_dtype_ok(dtype), LINA_ERR_BAD_ARG, "cross_entropy_rows: unknown dtype");
    LINA_DISPATCH_DTYPE(dtype, cross_entropy_rows_kernel<T_><<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
                                   (const T_ *)logits, ld, (const long long *)target, row_mask, loss, valid, M, Vn,
                                   ignore_index));
    LINA_LAUNCH_OK("cross_entropy_rows_kernel");
    return LINA_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm for the autocast TRAINING path: fp32 residual stream in, activation-dtype (bf16) output, statistics saved.
// Under torch autocast nn.LayerNorm returns fp32 and every consuming Linear (five in the GLA mixer) casts it to bf16 again:
// 34 bytes per element of traffic for norm1 instead of the 6 of this kernel; values reaching the GEMMs are identical.
namespace {

template <typename TO, int CH>
__global__ void __launch_bounds__(256)
layernorm_f32in_fwd_kernel(const float *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta,
                           TO *__restrict__ y, float *__restrict__ mean_out, float *__restrict__ rstd_out, int M, int N,
                           float eps) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const int nch = N / 4;                         // float4 chunks; host guarantees N % 4 == 0, nch <= 32 * CH
    const float4 *xr = reinterpret_cast<const float4 *>(x + (size_t)row * N);
    float4 v[CH];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int ch = lane + c * 32;
        v[c] = ch < nch ? xr[ch] : make_float4(0.f, 0.f, 0.f, 0.f);
        s += (v[c].x + v[c].y) + (v[c].z + v[c].w);
    }
    const float mean = warp_sum(s) / (float)N;
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        if (lane + c * 32 < nch) {
            const float a = v[c].x - mean, b = v[c].y - mean, d = v[c].z - mean, e = v[c].w - mean;
            ss += (a * a + b * b) + (d * d + e * e);
        }
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)N + eps);
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
    TO *yr = y + (size_t)row * N;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int ch = lane + c * 32;
        if (ch < nch) {
            const float4 g = reinterpret_cast<const float4 *>(gamma)[ch], b = reinterpret_cast<const float4 *>(beta)[ch];
            const float o0 = (v[c].x - mean) * rstd * g.x + b.x, o1 = (v[c].y - mean) * rstd * g.y + b.y;
            const float o2 = (v[c].z - mean) * rstd * g.z + b.z, o3 = (v[c].w - mean) * rstd * g.w + b.w;
            yr[ch * 4 + 0] = from_f<TO>(o0); yr[ch * 4 + 1] = from_f<TO>(o1);
            yr[ch * 4 + 2] = from_f<TO>(o2); yr[ch * 4 + 3] = from_f<TO>(o3);
        }
    }
}

// dx = rstd * (dxh - mean(dxh) - xh * mean(dxh * xh)), dxh = dy * gamma, xh = (x - mean) * rstd ; dgamma += dy * xh ; dbeta += dy
// Persistent grid (2 blocks per SM), rows strided over the warps; dgamma / dbeta reduced in registers, then per block in
// shared memory, then one atomic per column per block.
template <typename TO, int CH>
__global__ void __launch_bounds__(256)
layernorm_f32in_bwd_kernel(const float *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ mean_in,
                           const float *__restrict__ rstd_in, const TO *__restrict__ dy, float *__restrict__ dx,
                           float *__restrict__ dgamma, float *__restrict__ dbeta, int M, int N) {
    __shared__ float red[8][32 * CH * 4 + 1];
    const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nch = N / 4;
    float dg[CH][4], db[CH][4];
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) { dg[c][i] = 0.f; db[c][i] = 0.f; }
    for (int row = blockIdx.x * 8 + wl; row < M; row += gridDim.x * 8) {
        const float mean = mean_in[row], rstd = rstd_in[row];
        const float4 *xr = reinterpret_cast<const float4 *>(x + (size_t)row * N);
        const TO *dyr = dy + (size_t)row * N;
        float4 xv[CH];
        float dyv[CH][4];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int ch = lane + c * 32;
            if (ch < nch) {
                xv[c] = xr[ch];
#pragma unroll
                for (int i = 0; i < 4; ++i) dyv[c][i] = to_f(dyr[ch * 4 + i]);
            }
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int ch = lane + c * 32;
            if (ch < nch) {
                const float4 g = reinterpret_cast<const float4 *>(gamma)[ch];
                const float xs[4] = {xv[c].x, xv[c].y, xv[c].z, xv[c].w}, gs[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float xh = (xs[i] - mean) * rstd, dxh = dyv[c][i] * gs[i];
                    s1 += dxh;
                    s2 = fmaf(dxh, xh, s2);
                    dg[c][i] = fmaf(dyv[c][i], xh, dg[c][i]);
                    db[c][i] += dyv[c][i];
                }
            }
        }
        s1 = warp_sum(s1) / (float)N;
        s2 = warp_sum(s2) / (float)N;
        float4 *dxr = reinterpret_cast<float4 *>(dx + (size_t)row * N);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int ch = lane + c * 32;
            if (ch < nch) {
                const float4 g = reinterpret_cast<const float4 *>(gamma)[ch];
                const float xs[4] = {xv[c].x, xv[c].y, xv[c].z, xv[c].w}, gs[4] = {g.x, g.y, g.z, g.w};
                float o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float xh = (xs[i] - mean) * rstd;
                    o[i] = rstd * (dyv[c][i] * gs[i] - s1 - xh * s2);
                }
                dxr[ch] = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
    }
    // block reduction of dgamma, then of dbeta, through one shared buffer; one atomic per column per block
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
        __syncthreads();
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
            for (int i = 0; i < 4; ++i) red[wl][(lane + c * 32) * 4 + i] = which == 0 ? dg[c][i] : db[c][i];
        __syncthreads();
        float *dst = which == 0 ? dgamma : dbeta;
        for (int col = threadIdx.x; col < N; col += 256) {
            float a = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) a += red[w8][col];
            atomicAdd(&dst[col], a);
        }
    }
}

}  // namespace

extern "C" int lina_layernorm_f32in_fwd(const float *x, const float *gamma, const float *beta, void *y, float *mean,
                                        float *rstd, int M, int N, float eps, int out_dtype, void *stream) {
    LINA_REQUIRE(x && gamma && beta && y && mean && rstd && M > 0 && N > 0, LINA_ERR_BAD_ARG, "layernorm_f32in_fwd: bad argument");
    LINA_REQUIRE(lina_dtype_ok(out_dtype), LINA_ERR_BAD_ARG, "layernorm_f32in_fwd: unknown dtype");
    LINA_REQUIRE(N % 4 == 0 && N / 4 <= 32 * 8 && (uintptr_t)x % 16 == 0 && (uintptr_t)gamma % 16 == 0 &&
                     (uintptr_t)beta % 16 == 0, LINA_ERR_UNSUPPORTED,
                 "layernorm_f32in_fwd: N=%d must be a multiple of 4, <= 1024, tensors 16-byte aligned", N);
    LINA_DISPATCH_DTYPE(out_dtype, layernorm_f32in_fwd_kernel<T_, 8><<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
                                       x, gamma, beta, (T_ *)y, mean, rstd, M, N, eps));
    LINA_LAUNCH_OK("layernorm_f32in_fwd_kernel");
    return LINA_OK;
}

extern "C" int lina_layernorm_f32in_bwd(const float *x, const float *gamma, const float *mean, const float *rstd,
                                        const void *dy, float *dx, float *dgamma, float *dbeta, int M, int N, int dy_dtype,
                                        void *stream) {
    LINA_REQUIRE(x && gamma && mean && rstd && dy && dx && dgamma && dbeta && M > 0 && N > 0, LINA_ERR_BAD_ARG,
                 "layernorm_f32in_bwd: bad argument");
    LINA_REQUIRE(lina_dtype_ok(dy_dtype), LINA_ERR_BAD_ARG, "layernorm_f32in_bwd: unknown dtype");
    LINA_REQUIRE(N % 4 == 0 && N / 4 <= 32 * 8 && (uintptr_t)x % 16 == 0 && (uintptr_t)dx % 16 == 0 &&
                     (uintptr_t)gamma % 16 == 0, LINA_ERR_UNSUPPORTED,
                 "layernorm_f32in_bwd: N=%d must be a multiple of 4, <= 1024, tensors 16-byte aligned", N);
    int nblk = (M + 7) / 8;
    if (nblk > 148) nblk = 148;                 // persistent: one 8-warp block per SM (156 registers), rows strided
    LINA_DISPATCH_DTYPE(dy_dtype, layernorm_f32in_bwd_kernel<T_, 8><<<nblk, 256, 0, (cudaStream_t)stream>>>(
                                      x, gamma, mean, rstd, (const T_ *)dy, dx, dgamma, dbeta, M, N));
    LINA_LAUNCH_OK("layernorm_f32in_bwd_kernel");
    return LINA_OK;
}
