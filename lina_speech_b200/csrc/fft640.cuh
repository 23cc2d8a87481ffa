// ISTFT head for n_fft = 1280 (the shipped WavTokenizer, DEC/heads.py:42-67 + DEC/spectral_ops.py:57-58): one WARP per frame.
//   polar -> X[0..640]                                   (mag = min(exp(m), 100), X = mag e^{ip})
//   Z[k] = (X[k] + conj X[640-k]) + i e^{i pi k/640} (X[k] - conj X[640-k])      (1280-point C2R as a 640-point complex FFT)
//   640 = 10 * 4 * 4 * 4 Stockham autosort stages, every stage an exact multiple of 32 butterflies (2, 5, 5, 5 per lane),
//   compile-time radices / strides, per-stage twiddle tables laid out so that a warp's lookups are bank-conflict free,
//   x[2n] = Re z[n], x[2n+1] = Im z[n], times window / 1280, written as float2.
// The phases are plain functions of (lane, pointers) separated by __syncwarp() in the kernel; within a phase every lane reads
// only what the previous phase wrote and writes its own cells.  That makes them runnable on the HOST one lane after the
// other: tests/test_host.py compiles this header with g++ and checks the whole index arithmetic against numpy's irfft
// without a device (the CUDA kernel in codec.cu adds only the launch geometry and the __syncwarp()s).
#pragma once
#include <math.h>
#if defined(__CUDACC__)
#define FFT640_HD __host__ __device__ __forceinline__
#else
#define FFT640_HD inline
#endif

namespace fft640 {

constexpr int M = 640, N = 1280, NBINS = M + 1;
// per-CTA tables (floats): pre[2][640] | st2[2][30] | st3[2][120] | st4[2][480] | win[1280]
constexpr int T_PRE = 0, T_ST2 = T_PRE + 2 * M, T_ST3 = T_ST2 + 2 * 30, T_ST4 = T_ST3 + 2 * 120, T_WIN = T_ST4 + 2 * 480;
constexpr int TABLE_FLOATS = T_WIN + N;
// per-warp scratch (floats): re0[641] im0[641] re1[640] im1[640]
constexpr int WARP_FLOATS = 2 * NBINS + 2 * M;

FFT640_HD void sincospi_(float x, float *s, float *c) {
#if defined(__CUDA_ARCH__)
    sincospif(x, s, c);
#else
    *s = (float)sin(3.14159265358979323846 * (double)x);
    *c = (float)cos(3.14159265358979323846 * (double)x);
#endif
}

// table entry i of TABLE_FLOATS (any thread may fill any entry)
FFT640_HD float table_entry(int i, const float *window) {
    float s, c;
    if (i >= T_WIN) return window[i - T_WIN] * (1.f / (float)N);
    int base, Ns;
    if (i >= T_ST4) { base = T_ST4; Ns = 160; }
    else if (i >= T_ST3) { base = T_ST3; Ns = 40; }
    else if (i >= T_ST2) { base = T_ST2; Ns = 10; }
    else {                                              // pre[k] = e^{i pi k / 640}
        const int k = i % M;
        sincospi_((float)k / (float)M, &s, &c);
        return i < M ? c : s;
    }
    // stage table: entry (kk, t-1), t = 1..3, holds e^{2 pi i t kk / (4 Ns)}; cos plane then sin plane
    const int n = 3 * Ns, e = (i - base) % n, kk = e / 3, t = e % 3 + 1;
    sincospi_((float)(t * kk) / (float)(2 * Ns), &s, &c);
    return (i - base) < n ? c : s;
}

// phase 0: polar.  hf = one row of the head output: [641 log-magnitudes | 641 phases]
FFT640_HD void phase_polar(int lane, const float *hf, float *re0, float *im0) {
    for (int k = lane; k < NBINS; k += 32) {
        const float mag = fminf(expf(hf[k]), 100.f);
        float s, c;
        sincosf(hf[NBINS + k], &s, &c);
        re0[k] = mag * c;
        im0[k] = (k == 0 || k == M) ? 0.f : mag * s;    // C2R ignores the imaginary part of DC / Nyquist
    }
}

// 5-point DFT, kernel e^{+2 pi i t u / 5}, in place on (xr, xi)
FFT640_HD void dft5(float *xr, float *xi) {
    const float c1 = 0.30901699437494745f, c2 = -0.80901699437494745f, s1 = 0.95105651629515353f, s2 = 0.58778525229247314f;
    const float a1r = xr[1] + xr[4], a1i = xi[1] + xi[4], a2r = xr[2] + xr[3], a2i = xi[2] + xi[3];
    const float b1r = xr[1] - xr[4], b1i = xi[1] - xi[4], b2r = xr[2] - xr[3], b2i = xi[2] - xi[3];
    const float t1r = xr[0] + c1 * a1r + c2 * a2r, t1i = xi[0] + c1 * a1i + c2 * a2i;
    const float t2r = xr[0] + c2 * a1r + c1 * a2r, t2i = xi[0] + c2 * a1i + c1 * a2i;
    const float u1r = s1 * b1r + s2 * b2r, u1i = s1 * b1i + s2 * b2i;
    const float u2r = s2 * b1r - s1 * b2r, u2i = s2 * b1i - s1 * b2i;
    xr[0] += a1r + a2r;  xi[0] += a1i + a2i;
    xr[1] = t1r - u1i;   xi[1] = t1i + u1r;             // t1 + i u1
    xr[4] = t1r + u1i;   xi[4] = t1i - u1r;
    xr[2] = t2r - u2i;   xi[2] = t2i + u2r;
    xr[3] = t2r + u2i;   xi[3] = t2i - u2r;
}

// phase 1: radix-10 stage (Ns = 1, no twiddles) reading Z[k] formed on the fly from X; butterfly j reads k = j + 64 t and
// writes dst[10 j + u]
FFT640_HD void phase_r10(int lane, const float *re0, const float *im0, const float *tab, float *re1, float *im1) {
    const float *prr = tab + T_PRE, *pri = prr + M;
    for (int j = lane; j < 64; j += 32) {
        float er[5], ei[5], orr[5], oi[5];
#pragma unroll
        for (int t = 0; t < 10; ++t) {
            const int k = j + 64 * t;
            const float ar = re0[k], ai = im0[k], br = re0[M - k], bi = -im0[M - k];
            const float sr = ar + br, si = ai + bi, dr = ar - br, di = ai - bi;
            const float wr = prr[k], wi = pri[k];
            const float tr = wr * dr - wi * di, ti = wr * di + wi * dr;      // w^k * dif;  times i -> (-ti, tr)
            const float zr = sr - ti, zi = si + tr;
            if (t & 1) { orr[t >> 1] = zr; oi[t >> 1] = zi; } else { er[t >> 1] = zr; ei[t >> 1] = zi; }
        }
        dft5(er, ei);
        dft5(orr, oi);
        // X[u] = E[u] + w10^u O[u], X[u+5] = E[u] - w10^u O[u],  w10 = e^{2 pi i / 10}
        const float wr[5] = {1.f, 0.80901699437494745f, 0.30901699437494745f, -0.30901699437494745f, -0.80901699437494745f};
        const float wi[5] = {0.f, 0.58778525229247314f, 0.95105651629515353f, 0.95105651629515353f, 0.58778525229247314f};
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            const float pr = wr[u] * orr[u] - wi[u] * oi[u], pi = wr[u] * oi[u] + wi[u] * orr[u];
            re1[10 * j + u] = er[u] + pr;      im1[10 * j + u] = ei[u] + pi;
            re1[10 * j + u + 5] = er[u] - pr;  im1[10 * j + u + 5] = ei[u] - pi;
        }
    }
}

// one radix-4 butterfly of the stage with Ns finished points: loads src[j + 160 t] * tw(t, kk), returns the 4 outputs
template <int Ns>
FFT640_HD void r4_butterfly(int j, const float *sr, const float *si, const float *tab, int tab_off, float *yr, float *yi, int *j0) {
    const float *twr = tab + tab_off, *twi = twr + 3 * Ns;
    const int kk = j % Ns;
    float vr[4], vi[4];
    vr[0] = sr[j];
    vi[0] = si[j];
#pragma unroll
    for (int t = 1; t < 4; ++t) {
        const float xr = sr[j + 160 * t], xi = si[j + 160 * t];
        const float wr = twr[3 * kk + t - 1], wi = twi[3 * kk + t - 1];
        vr[t] = xr * wr - xi * wi;
        vi[t] = xr * wi + xi * wr;
    }
    const float pr = vr[0] + vr[2], pi = vi[0] + vi[2], qr = vr[0] - vr[2], qi = vi[0] - vi[2];
    const float rr = vr[1] + vr[3], ri = vi[1] + vi[3];
    const float dr = vr[1] - vr[3], di = vi[1] - vi[3];               // times i -> (-di, dr)
    yr[0] = pr + rr;  yi[0] = pi + ri;
    yr[1] = qr - di;  yi[1] = qi + dr;
    yr[2] = pr - rr;  yi[2] = pi - ri;
    yr[3] = qr + di;  yi[3] = qi - dr;
    *j0 = (j / Ns) * (4 * Ns) + kk;
}

// phases 2, 3: radix-4 stage from (sr, si) to (dr, di)
template <int Ns>
FFT640_HD void phase_r4(int lane, const float *sr, const float *si, const float *tab, int tab_off, float *dr, float *di) {
    for (int j = lane; j < 160; j += 32) {
        float yr[4], yi[4];
        int j0;
        r4_butterfly<Ns>(j, sr, si, tab, tab_off, yr, yi, &j0);
#pragma unroll
        for (int u = 0; u < 4; ++u) { dr[j0 + u * Ns] = yr[u]; di[j0 + u * Ns] = yi[u]; }
    }
}

// phase 4: last radix-4 stage (Ns = 160) straight to the windowed frame: z[n] -> fr[2n], fr[2n+1]
FFT640_HD void phase_r4_out(int lane, const float *sr, const float *si, const float *tab, float *fr) {
    const float *win = tab + T_WIN;
    for (int j = lane; j < 160; j += 32) {
        float yr[4], yi[4];
        int j0;
        r4_butterfly<160>(j, sr, si, tab, T_ST4, yr, yi, &j0);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int n = j0 + u * 160;
#if defined(__CUDA_ARCH__)
            *reinterpret_cast<float2 *>(fr + 2 * n) = make_float2(yr[u] * win[2 * n], yi[u] * win[2 * n + 1]);
#else
            fr[2 * n] = yr[u] * win[2 * n];
            fr[2 * n + 1] = yi[u] * win[2 * n + 1];
#endif
        }
    }
}

}  // namespace fft640
