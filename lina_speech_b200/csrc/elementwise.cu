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
            const uint4 gr = *reinterpret_cast<const uint4 *>(gamma + (size_t)ch * VEC);
            const uint4 br = *reinterpret_cast<const uint4 *>(beta + (size_t)ch * VEC);
            const T *ge = reinterpret_cast<const T *>(&gr), *be = reinterpret_cast<const T *>(&br);
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
