// Packed fp32x2 helpers for the bf16 streaming kernels (sm_100 FFMA2 / FMUL2: two lanes per issue slot).
// At the full HBM rate a bf16 read+write kernel has only ~22 issue slots per element on B200 (23 B/clk/SM against 128
// lanes/clk/SM); the scalar fp32 short conv used ~16 of them and sat at 0.43-0.59 of the copy bandwidth.
#pragma once
#include "common.cuh"

static __device__ __forceinline__ float2 bf2_to_f2(uint32_t w) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
static __device__ __forceinline__ uint32_t f2_to_bf2(float2 a) {     // (a.x -> low half, a.y -> high half), round to nearest even
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a.y), "f"(a.x));
    return r;
}
static __device__ __forceinline__ float2 silu2(float2 a) {          // a * sigmoid(a) = h * tanh(h) + h, h = a / 2
    const float2 h = __fmul2_rn(a, make_float2(0.5f, 0.5f));
    const float2 t = make_float2(tanh_approx_(h.x), tanh_approx_(h.y));
    return __ffma2_rn(h, t, h);
}
// taps of the channel pair (c, c+1) from 16 bytes = w[c][0..3], w[c+1][0..3]  ->  wt[j] = (w[c][j], w[c+1][j])
static __device__ __forceinline__ void load_taps2(const bf16 *wp, float2 (&wt)[4]) {
    const uint4 r = *reinterpret_cast<const uint4 *>(wp);
    const float2 a0 = bf2_to_f2(r.x), a1 = bf2_to_f2(r.y), b0 = bf2_to_f2(r.z), b1 = bf2_to_f2(r.w);
    wt[0] = make_float2(a0.x, b0.x); wt[1] = make_float2(a0.y, b0.y);
    wt[2] = make_float2(a1.x, b1.x); wt[3] = make_float2(a1.y, b1.y);
}
// one causal-conv output for a channel pair + window shift
static __device__ __forceinline__ float2 conv4_2(float2 (&win)[3], float2 x, const float2 (&wt)[4]) {
    float2 acc = __fmul2_rn(win[0], wt[0]);
    acc = __ffma2_rn(win[1], wt[1], acc);
    acc = __ffma2_rn(win[2], wt[2], acc);
    acc = __ffma2_rn(x, wt[3], acc);
    win[0] = win[1]; win[1] = win[2]; win[2] = x;
    return acc;
}

