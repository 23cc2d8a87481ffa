// Shared helpers for the liblina_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/lina_b200.h"

typedef __nv_bfloat16 bf16;

// ---- per-thread error string ------------------------------------------------------------------
void lina_set_error(const char *fmt, ...);

#define LINA_REQUIRE(cond, code, ...)                  \
    do {                                               \
        if (!(cond)) {                                 \
            lina_set_error(__VA_ARGS__);               \
            return (code);                             \
        }                                              \
    } while (0)

#define LINA_CUDA_OK(call)                                                                   \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            lina_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return LINA_ERR_CUDA;                                                            \
        }                                                                                    \
    } while (0)

#define LINA_LAUNCH_OK(name)                                                                 \
    do {                                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess) {                                                            \
            lina_set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));        \
            return LINA_ERR_CUDA;                                                            \
        }                                                                                    \
    } while (0)

// cudaFuncSetAttribute applies to the CURRENT device only: a call site remembers per device (bit = ordinal mod 64) whether
// it has configured its kernel.  Usage: static thread_local uint64_t done = 0; if (lina_first_use_on_device(&done)) {...}
static inline bool lina_first_use_on_device(uint64_t *done_mask) {
    int dev = 0;
    (void)cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (dev & 63);
    if (*done_mask & bit) return false;
    *done_mask |= bit;
    return true;
}

// ---- dtype conversion -------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T x);
template <> __device__ __forceinline__ float to_f<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 x) { return __bfloat162float(x); }
template <> __device__ __forceinline__ float to_f<__half>(__half x) { return __half2float(x); }

template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }
template <> __device__ __forceinline__ __half from_f<__half>(float x) { return __float2half_rn(x); }

// element of run-time dtype (used outside hot loops: initial states)
__device__ __forceinline__ float load_dyn(const void *p, int dtype, size_t i) {
    if (dtype == LINA_F32) return ((const float *)p)[i];
    if (dtype == LINA_BF16) return __bfloat162float(((const bf16 *)p)[i]);
    return __half2float(((const __half *)p)[i]);
}
__device__ __forceinline__ void store_dyn(void *p, int dtype, size_t i, float x) {
    if (dtype == LINA_F32) ((float *)p)[i] = x;
    else if (dtype == LINA_BF16) ((bf16 *)p)[i] = __float2bfloat16_rn(x);
    else ((__half *)p)[i] = __float2half_rn(x);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
// one MUFU op instead of two (ex2 + rcp): sigmoid(x) = 0.5 tanh(x/2) + 0.5 with tanh.approx (rel. error ~2^-11).
// Used only where the result is rounded to a 16-bit type right away; fp32 I/O keeps the exact form.
__device__ __forceinline__ float tanh_approx_(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <typename T> __device__ __forceinline__ float sigmoid_io(float x) {
    if (sizeof(T) == 2) return fmaf(0.5f, tanh_approx_(0.5f * x), 0.5f);
    return sigmoidf_(x);
}
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }
// log(sigmoid(x)) = min(x,0) - log1p(exp(-|x|))   (stable for both tails)
__device__ __forceinline__ float logsigmoidf_(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

static inline int lina_dtype_ok(int dt) { return dt == LINA_F32 || dt == LINA_BF16 || dt == LINA_F16; }
static inline size_t lina_dtype_size(int dt) { return dt == LINA_F32 ? 4 : 2; }

// dispatch a templated launcher on the activation dtype
#define LINA_DISPATCH_DTYPE(dt, ...)                           \
    switch (dt) {                                              \
        case LINA_F32: { typedef float T_; __VA_ARGS__; } break;   \
        case LINA_BF16: { typedef bf16 T_; __VA_ARGS__; } break;   \
        case LINA_F16: { typedef __half T_; __VA_ARGS__; } break;  \
        default: lina_set_error("unknown dtype %d", dt); return LINA_ERR_BAD_ARG; \
    }
