// WavTokenizer decode tail: the HBM-bound stages between the backbone's dense GEMMs / convolutions.
//
// Reference (DEC/ = 3rdparty/decoder/): codes_to_features DEC/pretrained.py:209-239; GroupNorm+swish
// DEC/models.py:10-16,58-70; ConvNeXtBlock DEC/modules.py:43-60; AdaLayerNorm DEC/modules.py:81-86;
// final LayerNorm DEC/models.py:234; ISTFTHead DEC/heads.py:53-67; ISTFT("same") DEC/spectral_ops.py:33-75.
// Everything is fp32 like the reference.  Layout notes:
//   backbone activations are [B,C,L] (channel-major) in the reference; the ConvNeXt MLP wants [B,L,C].
//   dwconv_adaln fuses the depthwise conv, the transpose and the (Ada)LayerNorm in one pass over x;
//   scale_residual_t fuses gamma*h, the transpose back and the residual add.
#include "common.cuh"
#include "fft640.cuh"

extern int g_lina_variant[16];

namespace {

// ------------------------------------------------------------------------------------------------
// codes -> features: out[b,c,l] = sum_k codebooks[k*bins + codes[k,b,l], c]   (32x32 smem transpose)
__global__ void __launch_bounds__(256)
codes_to_features_kernel(const int64_t *__restrict__ codes, const float *__restrict__ books,
                         float *__restrict__ out, int Kq, int B, int L, int bins, int C) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    for (int i = ty; i < 32; i += 8) {                               // i = l within tile, tx = c
        const int l = l0 + i, c = c0 + tx;
        float acc = 0.f;
        if (l < L && c < C) {
            for (int k = 0; k < Kq; ++k) {
                long long id = codes[((size_t)k * B + b) * L + l];
                id = id < 0 ? 0 : (id >= bins ? bins - 1 : id);
                acc += books[((size_t)k * bins + id) * C + c];
            }
        }
        tile[i][tx] = acc;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {                               // i = c within tile, tx = l
        const int c = c0 + i, l = l0 + tx;
        if (c < C && l < L) out[((size_t)b * C + c) * L + l] = tile[tx][i];
    }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm (+swish): one CTA per (b, group); the group's channels are one contiguous run of cpg*L floats.
__global__ void __launch_bounds__(256)
groupnorm_swish_kernel(const float *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta,
                       float *__restrict__ y, int C, int L, int groups, float eps, int swish) {
    __shared__ double red[2][8];
    __shared__ float stat[2];
    const int bg = blockIdx.x, g = bg % groups;
    const int cpg = C / groups;
    const size_t n = (size_t)cpg * L;
    const float *xp = x + (size_t)bg * n;
    float *yp = y + (size_t)bg * n;
    float s = 0.f, ss = 0.f;
    for (size_t i = threadIdx.x; i < n; i += 256) { const float v = xp[i]; s += v; ss = fmaf(v, v, ss); }
    double ds = warp_sum(s), dss = warp_sum(ss);
    // (warp_sum is float; widen across warps)
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ds; red[1][threadIdx.x >> 5] = dss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b2 = 0;
        for (int w = 0; w < 8; ++w) { a += red[0][w]; b2 += red[1][w]; }
        const double mean = a / (double)n;
        double var = b2 / (double)n - mean * mean;
        if (var < 0) var = 0;
        stat[0] = (float)mean;
        stat[1] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    const float mean = stat[0], rstd = stat[1];
    for (size_t i = threadIdx.x; i < n; i += 256) {
        const int c = g * cpg + (int)(i / L);
        float v = (xp[i] - mean) * rstd * gamma[c] + beta[c];
        if (swish) v = v * sigmoidf_(v);
        yp[i] = v;
    }
}


// Register-resident variant: one CTA per (b, group), NT threads, the group's cpg*L floats (a contiguous run) are loaded
// ONCE as float4 (all loads issued before the first use), reduced, normalised from registers and stored: one read + one
// write of the tensor, >= 256 B in flight per thread.  Used when the run fits PER float4 per thread and is 16-byte aligned.
template <int NT, int PER>
__global__ void __launch_bounds__(NT)
groupnorm_swish_reg_kernel(const float *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta,
                           float *__restrict__ y, int C, int L, int groups, float eps, int swish) {
    __shared__ float red[2][NT / 32];
    __shared__ float stat[2];
    const int bg = blockIdx.x, g = bg % groups;
    const int cpg = C / groups;
    const int n4 = cpg * L / 4;                              // host guarantees (cpg * L) % 4 == 0
    const float4 *xp = reinterpret_cast<const float4 *>(x + (size_t)bg * cpg * L);
    float4 *yp = reinterpret_cast<float4 *>(y + (size_t)bg * cpg * L);
    float4 v[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int j = threadIdx.x + i * NT;
        v[i] = j < n4 ? xp[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0;
        for (int w = 0; w < NT / 32; ++w) a += red[0][w];
        stat[0] = (float)(a / (double)(cpg * L));
    }
    __syncthreads();
    const float mean = stat[0];
    float ss = 0.f;                                          // centred second moment: no cancellation
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        if (threadIdx.x + i * NT < n4) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            ss += (a * a + b * b) + (c * c + d * d);
        }
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0;
        for (int w = 0; w < NT / 32; ++w) a += red[1][w];
        stat[1] = (float)(1.0 / sqrt(a / (double)(cpg * L) + (double)eps));
    }
    __syncthreads();
    const float rstd = stat[1];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int j = threadIdx.x + i * NT;
        if (j < n4) {
            // one division per float4; its 4 elements straddle a channel boundary only when L % 4 != 0
            const int e0 = j * 4;
            const int cl = e0 / L, rem = e0 - cl * L;
            const int c0 = g * cpg + cl;
            const int c1 = min(c0 + 1, C - 1);
            const float ga0 = gamma[c0] * rstd, be0 = beta[c0] - mean * ga0;       // (x - mean) rstd gamma + beta = x A + B
            const float ga1 = gamma[c1] * rstd, be1 = beta[c1] - mean * ga1;
            float o[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool nxt = rem + q >= L;
                float t = fmaf(o[q], nxt ? ga1 : ga0, nxt ? be1 : be0);
                if (swish) t = __fdividef(t, 1.f + __expf(-t));
                o[q] = t;
            }
            yp[j] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// depthwise conv k=7 (optional) + transpose [B,C,L]->[B,L,C] + LayerNorm(no affine)*scale+shift.
// One CTA = one (batch, tile of TL time steps) x ALL channels.  Every channel row of the tile is one 128-byte
// warp load (32 floats: TL = 26 outputs + 6 halo with the conv, TL = 32 without), all rows are requested before
// any is used, the conv runs out of shared memory, the transposed tile is normalised by one warp per time step
// and written as full contiguous rows of [B,L,C].
constexpr int DW_THREADS = 512;
template <bool CONV>
__global__ void __launch_bounds__(DW_THREADS)
dwconv_adaln_kernel(const float *__restrict__ x, const float *__restrict__ dw_w, const float *__restrict__ dw_b,
                    const float *__restrict__ scale, const float *__restrict__ shift, float *__restrict__ y,
                    int C, int L, float eps) {
    constexpr int TL = CONV ? 26 : 32;
    extern __shared__ float smem[];
    const int CP = C + 1;
    float *outs = smem;                         // [TL][C + 1]   (transposed tile)
    float *ins = outs + TL * CP;                // [C][33]       (CONV only: raw rows incl. halo)
    float *wts = ins + (CONV ? C * 33 : 0);     // [C][8]        (CONV only: 7 taps + bias)
    const int b = blockIdx.y, l0 = blockIdx.x * TL;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float *xb = x + (size_t)b * C * L;
    if (CONV) {
        {   // 8 independent 128-byte row loads in flight per warp
            constexpr int NW = DW_THREADS / 32, U = 8;
            const int l = l0 - 3 + lane;
            const bool ok = l >= 0 && l < L;
            for (int c = warp; c < C; c += NW * U) {
                float v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) { const int cc = c + u * NW; v[u] = (ok && cc < C) ? xb[(size_t)cc * L + l] : 0.f; }
#pragma unroll
                for (int u = 0; u < U; ++u) { const int cc = c + u * NW; if (cc < C) ins[cc * 33 + lane] = v[u]; }
            }
        }
        for (int i = tid; i < C * 8; i += DW_THREADS) {
            const int c = i >> 3, j = i & 7;
            wts[i] = j < 7 ? dw_w[c * 7 + j] : (dw_b != nullptr ? dw_b[c] : 0.f);
        }
        __syncthreads();
        for (int e = tid; e < C * TL; e += DW_THREADS) {
            const int c = e / TL, lt = e - c * TL;
            const float *w = wts + c * 8, *in = ins + c * 33 + lt;
            float acc = w[7];
#pragma unroll
            for (int j = 0; j < 7; ++j) acc = fmaf(w[j], in[j], acc);
            outs[lt * CP + c] = acc;
        }
    } else {
        constexpr int NW = DW_THREADS / 32, U = 8;
        const int l = l0 + lane;
        for (int c = warp; c < C; c += NW * U) {
            float v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { const int cc = c + u * NW; v[u] = (l < L && cc < C) ? xb[(size_t)cc * L + l] : 0.f; }
#pragma unroll
            for (int u = 0; u < U; ++u) { const int cc = c + u * NW; if (cc < C) outs[lane * CP + cc] = v[u]; }
        }
    }
    __syncthreads();
    for (int lt = warp; lt < TL; lt += DW_THREADS / 32) {
        const int l = l0 + lt;
        if (l >= L) continue;
        const float *row = outs + lt * CP;
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += row[c];
        const float mean = warp_sum(s) / (float)C;
        float ss = 0.f;
        for (int c = lane; c < C; c += 32) { const float d = row[c] - mean; ss = fmaf(d, d, ss); }
        const float rstd = rsqrtf(warp_sum(ss) / (float)C + eps);
        float *yr = y + ((size_t)b * L + l) * C;
        for (int c = lane; c < C; c += 32) yr[c] = (row[c] - mean) * rstd * scale[c] + shift[c];
    }
}


// ------------------------------------------------------------------------------------------------
// Two-kernel form of dwconv + transpose + (Ada)LayerNorm (the production path; the single-kernel version above needs all C
// channels of a time tile in one CTA -> 205 KB of shared memory, one CTA per SM, phases serialised: 0.15 of the copy rate).
//   A: CTA = (b, 64 channels, 64 time steps): coalesced row loads (+-3 halo), 7-tap conv out of shared memory, transposed
//      store of the UN-normalised tile into y[b, l, c0..c0+63], per-row partial statistics (count, mean, M2 of the 64
//      channels) into ws[b, l, cblk, 2]  (35 KB smem, 6 CTAs per SM);
//   B: one warp per (b, l) row: Chan-merges the partials in a fixed order (deterministic, no cancellation), normalises the
//      row in place with float4 accesses.  The intermediate (73.7 MB at the shipped size) mostly lives in the 126 MB L2.
constexpr int DT_C = 64, DT_L = 64;
template <bool CONV>
__global__ void __launch_bounds__(256)
dwconv_t_stats_kernel(const float *__restrict__ x, const float *__restrict__ dw_w, const float *__restrict__ dw_b,
                      float *__restrict__ y, float *__restrict__ ws, int C, int L, int nblk) {
    __shared__ float ins[DT_C][DT_L + 9];                 // x[c][l0-3 .. l0+DT_L+3]  (CONV) or x[c][l0 .. l0+DT_L); odd stride: no bank conflicts
    __shared__ float outs[DT_L][DT_C + 1];                // conv output, transposed
    const int b = blockIdx.z, c0 = blockIdx.y * DT_C, l0 = blockIdx.x * DT_L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int HALO = CONV ? 3 : 0, ROW = DT_L + 2 * HALO;
    const float *xb = x + ((size_t)b * C + c0) * L;
    {   // every warp loads 8 channel rows; all its loads are issued before the first shared store
        float v[8][3];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = warp + u * 8;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int i = lane + k * 32, l = l0 - HALO + i;
                v[u][k] = (i < ROW && c0 + c < C && l >= 0 && l < L) ? xb[(size_t)c * L + l] : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int i = lane + k * 32;
                if (i < ROW) ins[warp + u * 8][i] = v[u][k];
            }
    }
    __syncthreads();
    {   // thread = channel (tid & 63) x 16 consecutive time steps, sliding 7-tap window (a warp = 32 channels, one time block)
        const int c = tid & 63, lb = (tid >> 6) * 16;
        if (CONV) {
            float w[7], bias = 0.f;
            const bool cok = c0 + c < C;
#pragma unroll
            for (int j = 0; j < 7; ++j) w[j] = cok ? dw_w[(size_t)(c0 + c) * 7 + j] : 0.f;
            if (cok && dw_b != nullptr) bias = dw_b[c0 + c];
            float win[22];
#pragma unroll
            for (int i = 0; i < 22; ++i) win[i] = ins[c][lb + i];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float acc = bias;
#pragma unroll
                for (int j = 0; j < 7; ++j) acc = fmaf(w[j], win[i + j], acc);
                outs[lb + i][c] = acc;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) outs[lb + i][c] = ins[c][lb + i];
        }
    }
    __syncthreads();
    const int nc = min(DT_C, C - c0);
    for (int r = warp; r < DT_L; r += 8) {               // one warp per time step: partial stats + transposed store
        const int l = l0 + r;
        if (l >= L) break;
        const float a0 = lane < nc ? outs[r][lane] : 0.f, a1 = lane + 32 < nc ? outs[r][lane + 32] : 0.f;
        const float mean = warp_sum(a0 + a1) / (float)nc;
        const float d0 = lane < nc ? a0 - mean : 0.f, d1 = lane + 32 < nc ? a1 - mean : 0.f;
        const float m2 = warp_sum(d0 * d0 + d1 * d1);
        float *yr = y + ((size_t)b * L + l) * C + c0;
        if (lane < nc) yr[lane] = a0;
        if (lane + 32 < nc) yr[lane + 32] = a1;
        if (lane == 0) {
            float *w2 = ws + (((size_t)b * L + l) * nblk + blockIdx.y) * 2;
            w2[0] = mean; w2[1] = m2;
        }
    }
}

__global__ void __launch_bounds__(256)
adaln_apply_kernel(float *__restrict__ y, const float *__restrict__ ws, const float *__restrict__ scale,
                   const float *__restrict__ shift, long long rows, int C, int nblk, float eps) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    // Chan merge of the per-block (mean, M2) partials, fixed order
    float mean = 0.f, m2 = 0.f, n = 0.f;
    for (int p = 0; p < nblk; ++p) {
        const float np = (float)min(DT_C, C - p * DT_C);
        const float mp = ws[((size_t)row * nblk + p) * 2], m2p = ws[((size_t)row * nblk + p) * 2 + 1];
        const float nt = n + np, delta = mp - mean;
        mean += delta * (np / nt);
        m2 += m2p + delta * delta * (n * np / nt);
        n = nt;
    }
    const float rstd = rsqrtf(m2 / (float)C + eps);
    float *yr = y + (size_t)row * C;
    if (C % 4 == 0 && ((uintptr_t)y % 16 == 0)) {
        const int n4 = C / 4;
        for (int j0 = 0; j0 < n4; j0 += 32 * 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = j0 + u * 32 + lane;
                if (j < n4) v[u] = reinterpret_cast<const float4 *>(yr)[j];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = j0 + u * 32 + lane;
                if (j < n4) {
                    const float4 sc = reinterpret_cast<const float4 *>(scale)[j], sh = reinterpret_cast<const float4 *>(shift)[j];
                    float4 o;
                    o.x = (v[u].x - mean) * rstd * sc.x + sh.x;
                    o.y = (v[u].y - mean) * rstd * sc.y + sh.y;
                    o.z = (v[u].z - mean) * rstd * sc.z + sh.z;
                    o.w = (v[u].w - mean) * rstd * sc.w + sh.w;
                    reinterpret_cast<float4 *>(yr)[j] = o;
                }
            }
        }
    } else {
        for (int c = lane; c < C; c += 32) yr[c] = (yr[c] - mean) * rstd * scale[c] + shift[c];
    }
}

// out[b,c,l] = res[b,c,l] + gamma[c] * h[b,l,c]
__global__ void __launch_bounds__(256)
scale_residual_t_kernel(const float *__restrict__ h, const float *__restrict__ gamma, const float *__restrict__ res,
                        float *__restrict__ out, int C, int L) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int l = l0 + i, c = c0 + tx;
        tile[i][tx] = (l < L && c < C) ? h[((size_t)b * L + l) * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, l = l0 + tx;
        if (c < C && l < L) {
            const size_t idx = ((size_t)b * C + c) * L + l;
            out[idx] = res[idx] + (gamma != nullptr ? gamma[c] : 1.f) * tile[tx][i];
        }
    }
}


// 64 x 64 tiles: h rows [l][c0..c0+63] are read as float4 (C % 4 == 0), transposed through shared memory, and res / out rows
// [c][l0..l0+63] are read / written as float2 (rows of [B,C,L] start on 8-byte boundaries when L is even; the shipped
// L = 750 is not a multiple of 4).  All loads of a thread are issued before the first use.
__global__ void __launch_bounds__(256)
scale_residual_t64_kernel(const float *__restrict__ h, const float *__restrict__ gamma, const float *__restrict__ res,
                          float *__restrict__ out, int C, int L) {
    __shared__ float tile[64][65];
    const int b = blockIdx.z, l0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    const int tid = threadIdx.x;
    float4 hv[4];                                          // 64 rows (l) x 16 float4 (c)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * 256, r = idx >> 4, q = idx & 15;
        const int l = l0 + r, c = c0 + q * 4;
        hv[i] = (l < L && c < C) ? *reinterpret_cast<const float4 *>(h + ((size_t)b * L + l) * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float2 rv[8];                                          // 64 rows (c) x 32 float2 (l)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = tid + i * 256, r = idx >> 5, q = idx & 31;
        const int c = c0 + r, l = l0 + q * 2;
        rv[i] = (c < C && l < L) ? *reinterpret_cast<const float2 *>(res + ((size_t)b * C + c) * L + l) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * 256, r = idx >> 4, q = idx & 15;
        tile[r][q * 4 + 0] = hv[i].x; tile[r][q * 4 + 1] = hv[i].y; tile[r][q * 4 + 2] = hv[i].z; tile[r][q * 4 + 3] = hv[i].w;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = tid + i * 256, r = idx >> 5, q = idx & 31;
        const int c = c0 + r, l = l0 + q * 2;
        if (c < C && l < L) {                              // L even: l + 1 < L too
            const float gm = gamma != nullptr ? gamma[c] : 1.f;
            *reinterpret_cast<float2 *>(out + ((size_t)b * C + c) * L + l) =
                make_float2(rv[i].x + gm * tile[q * 2][r], rv[i].y + gm * tile[q * 2 + 1][r]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ISTFT head: polar -> 1280-point inverse real FFT (as a 640-point complex Stockham FFT) -> window.
struct FftPlan { int M; int nstage; int radix[16]; };

struct cplx { float x, y; };
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.x + b.x, a.y + b.y}; }

__global__ void __launch_bounds__(128)
istft_frames_kernel(const float *__restrict__ h, long long ldh, const float *__restrict__ window, float *__restrict__ frames,
                    int nframes, FftPlan plan) {
    extern __shared__ float smem[];
    const int M = plan.M, N = 2 * M;
    cplx *tw = reinterpret_cast<cplx *>(smem);       // [M]   e^{+2 pi i m / M}
    cplx *pre = tw + M;                              // [M]   e^{+2 pi i k / N}
    cplx *bufa = pre + M;                            // [M+1] spectrum, then ping
    cplx *bufb = bufa + (M + 1);                     // [M]   pong
    const int tid = threadIdx.x;
    for (int m = tid; m < M; m += 128) {
        float s, c;
        sincospif(2.f * (float)m / (float)M, &s, &c);
        tw[m] = {c, s};
        sincospif((float)m / (float)M, &s, &c);      // 2 pi m / N = pi m / M
        pre[m] = {c, s};
    }
    for (int f = blockIdx.x; f < nframes; f += gridDim.x) {
        __syncthreads();
        const float *hf = h + (size_t)f * ldh;
        // X[k] = min(exp(m_k), 100) * (cos p_k + i sin p_k)        (DEC/heads.py:54-66)
        for (int k = tid; k <= M; k += 128) {
            const float mag = fminf(expf(hf[k]), 100.f);
            float s, c;
            sincosf(hf[M + 1 + k], &s, &c);
            cplx X = {mag * c, mag * s};
            if (k == 0 || k == M) X.y = 0.f;         // C2R ignores the imaginary part of DC / Nyquist
            bufa[k] = X;
        }
        __syncthreads();
        // Z[k] = (X[k] + conj X[M-k]) + i w^k (X[k] - conj X[M-k]),  w = e^{2 pi i / N}
        for (int k = tid; k < M; k += 128) {
            const cplx A = bufa[k], Bc = {bufa[M - k].x, -bufa[M - k].y};
            const cplx sum = cadd(A, Bc), dif = {A.x - Bc.x, A.y - Bc.y};
            const cplx t = cmul(pre[k], dif);        // w^k * dif ; times i -> (-t.y, t.x)
            bufb[k] = {sum.x - t.y, sum.y + t.x};
        }
        __syncthreads();
        cplx *a = bufb, *b = bufa;
        int Ns = 1;
        for (int st = 0; st < plan.nstage; ++st) {
            const int r = plan.radix[st], nb = M / r;
            for (int j = tid; j < nb; j += 128) {
                const int kk = j % Ns;
                cplx v[5];
                const int tstep = M / (Ns * r);
                for (int t = 0; t < r; ++t) {
                    cplx xin = a[j + t * nb];
                    v[t] = t == 0 ? xin : cmul(xin, tw[(t * kk * tstep) % M]);
                }
                const int j0 = (j / Ns) * Ns * r + kk;
                const int ustep = M / r;
                for (int u = 0; u < r; ++u) {
                    cplx acc = v[0];
                    for (int t = 1; t < r; ++t) acc = cadd(acc, cmul(v[t], tw[(t * u * ustep) % M]));
                    b[j0 + u * Ns] = acc;
                }
            }
            __syncthreads();
            cplx *tmp = a; a = b; b = tmp;
            Ns *= r;
        }
        // x[2n] = Re z[n], x[2n+1] = Im z[n], scaled 1/N, times the window  (spectral_ops.py:57-58)
        float *fr = frames + (size_t)f * N;
        const float inv = 1.f / (float)N;
        for (int n = tid; n < M; n += 128) {
            const cplx z = a[n];
            float2 o2 = make_float2(z.x * inv * window[2 * n], z.y * inv * window[2 * n + 1]);
            *reinterpret_cast<float2 *>(fr + 2 * n) = o2;
        }
    }
}

// n_fft = 1280: one warp per frame, fixed radices 10 * 4 * 4 * 4 (csrc/fft640.cuh holds the phases and their rationale).
// 8 warps per CTA, 97 KB of shared memory (15 KB of tables + 10 KB per warp), 2 CTAs per SM.
constexpr int F640_WARPS = 8;
constexpr size_t F640_SMEM = (size_t)(fft640::TABLE_FLOATS + F640_WARPS * fft640::WARP_FLOATS) * sizeof(float);

__global__ void __launch_bounds__(F640_WARPS * 32)
istft_frames_1280_kernel(const float *__restrict__ h, long long ldh, const float *__restrict__ window, float *__restrict__ frames,
                         int nframes) {
    using namespace fft640;
    extern __shared__ float smem[];
    float *tab = smem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < TABLE_FLOATS; i += F640_WARPS * 32) tab[i] = table_entry(i, window);
    __syncthreads();
    float *re0 = smem + TABLE_FLOATS + warp * WARP_FLOATS, *im0 = re0 + NBINS, *re1 = im0 + NBINS, *im1 = re1 + M;
    for (int f = blockIdx.x * F640_WARPS + warp; f < nframes; f += gridDim.x * F640_WARPS) {
        phase_polar(lane, h + (size_t)f * ldh, re0, im0);
        __syncwarp();
        phase_r10(lane, re0, im0, tab, re1, im1);
        __syncwarp();
        phase_r4<10>(lane, re1, im1, tab, T_ST2, re0, im0);
        __syncwarp();
        phase_r4<40>(lane, re0, im0, tab, T_ST3, re1, im1);
        __syncwarp();
        phase_r4_out(lane, re1, im1, tab, frames + (size_t)f * N);
        __syncwarp();                                   // the next frame's polar phase overwrites re0 / im0
    }
}

// overlap-add + trim + envelope normalisation (spectral_ops.py:60-73)
__global__ void __launch_bounds__(256)
istft_ola_kernel(const float *__restrict__ frames, const float *__restrict__ window, float *__restrict__ wav,
                 int L, int N, int hop) {
    const int b = blockIdx.y;
    const int s = blockIdx.x * 256 + threadIdx.x;
    const int out_len = L * hop;
    if (s >= out_len) return;
    const int pad = (N - hop) / 2;
    const int p = s + pad;
    int f_hi = p / hop;
    if (f_hi > L - 1) f_hi = L - 1;
    int f_lo = (p - N + hop) / hop;            // ceil((p - N + 1) / hop)
    if (p - N + 1 <= 0) f_lo = 0;
    float acc = 0.f, env = 0.f;
    for (int f = f_lo; f <= f_hi; ++f) {
        const int n = p - f * hop;
        if (n < 0 || n >= N) continue;
        acc += frames[((size_t)b * L + f) * N + n];
        const float w = window[n];
        env = fmaf(w, w, env);
    }
    wav[(size_t)b * out_len + s] = acc / env;
}

bool make_plan(int M, FftPlan &p) {
    p.M = M; p.nstage = 0;
    int m = M;
    const int cand[4] = {5, 4, 2, 3};
    for (int ci = 0; ci < 4; ++ci) {
        const int r = cand[ci];
        while (m % r == 0 && m > 1) {
            if (r == 2 && m % 4 == 0) break;     // (unreachable: 4s are removed first)
            if (p.nstage >= 16) return false;
            p.radix[p.nstage++] = r;
            m /= r;
        }
    }
    return m == 1;
}

}  // namespace

extern "C" int lina_codec_codes_to_features(const int64_t *codes, const float *codebooks, float *features, int Kq,
                                            int B, int L, int bins, int C, void *stream) {
    LINA_REQUIRE(codes && codebooks && features, LINA_ERR_BAD_ARG, "codes_to_features: null pointer");
    LINA_REQUIRE(Kq > 0 && B > 0 && L > 0 && bins > 0 && C > 0, LINA_ERR_BAD_ARG, "codes_to_features: bad size");
    LINA_REQUIRE(B <= 65535 && (C + 31) / 32 <= 65535, LINA_ERR_UNSUPPORTED, "codes_to_features: grid too large");
    dim3 grid((L + 31) / 32, (C + 31) / 32, B);
    codes_to_features_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(codes, codebooks, features, Kq, B, L, bins, C);
    LINA_LAUNCH_OK("codes_to_features_kernel");
    return LINA_OK;
}

extern "C" int lina_codec_groupnorm_swish(const float *x, const float *gamma, const float *beta, float *y, float *ws,
                                          int B, int C, int L, int groups, float eps, int swish, void *stream) {
    (void)ws;
    LINA_REQUIRE(x && gamma && beta && y, LINA_ERR_BAD_ARG, "groupnorm_swish: null pointer");
    LINA_REQUIRE(B > 0 && C > 0 && L > 0 && groups > 0 && C % groups == 0, LINA_ERR_BAD_ARG,
                 "groupnorm_swish: bad size (C=%d groups=%d)", C, groups);
    const long long run = (long long)(C / groups) * L;
    const bool al = ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && run % 4 == 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (al && run / 4 <= 512 * 12) {                 // <= 24576 floats per (b, group): 12 float4 per thread, 512 threads
        groupnorm_swish_reg_kernel<512, 12><<<B * groups, 512, 0, st>>>(x, gamma, beta, y, C, L, groups, eps, swish);
        LINA_LAUNCH_OK("groupnorm_swish_reg_kernel");
        return LINA_OK;
    }
    if (al && run / 4 <= 512 * 24) {                 // <= 49152 floats (L <= 2048 at 24 channels per group)
        groupnorm_swish_reg_kernel<512, 24><<<B * groups, 512, 0, st>>>(x, gamma, beta, y, C, L, groups, eps, swish);
        LINA_LAUNCH_OK("groupnorm_swish_reg_kernel");
        return LINA_OK;
    }
    groupnorm_swish_kernel<<<B * groups, 256, 0, st>>>(x, gamma, beta, y, C, L, groups, eps, swish);
    LINA_LAUNCH_OK("groupnorm_swish_kernel");
    return LINA_OK;
}

static int launch_dwconv_adaln(const float *x, const float *dw_w, const float *dw_b, const float *scale,
                               const float *shift, float *y, int B, int C, int L, float eps, cudaStream_t st) {
    LINA_REQUIRE(x && scale && shift && y, LINA_ERR_BAD_ARG, "dwconv_adaln: null pointer");
    LINA_REQUIRE(B > 0 && C > 0 && L > 0, LINA_ERR_BAD_ARG, "dwconv_adaln: bad size");
    LINA_REQUIRE(B <= 65535, LINA_ERR_UNSUPPORTED, "dwconv_adaln: B > 65535");
    const bool conv = dw_w != nullptr;
    const int TL = conv ? 26 : 32;
    const size_t smem = ((size_t)TL * (C + 1) + (conv ? (size_t)C * 41 : 0)) * sizeof(float);
    LINA_REQUIRE(smem <= 227 * 1024, LINA_ERR_UNSUPPORTED, "dwconv_adaln: C=%d too large for shared memory", C);
    static thread_local uint64_t configured[2] = {0, 0};
    if (lina_first_use_on_device(&configured[conv])) {            // once per device: allow the whole 227 KB
        if (conv) LINA_CUDA_OK(cudaFuncSetAttribute(dwconv_adaln_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        else LINA_CUDA_OK(cudaFuncSetAttribute(dwconv_adaln_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    dim3 grid((L + TL - 1) / TL, B);
    if (conv) dwconv_adaln_kernel<true><<<grid, DW_THREADS, smem, st>>>(x, dw_w, dw_b, scale, shift, y, C, L, eps);
    else dwconv_adaln_kernel<false><<<grid, DW_THREADS, smem, st>>>(x, dw_w, dw_b, scale, shift, y, C, L, eps);
    LINA_LAUNCH_OK("dwconv_adaln_kernel");
    return LINA_OK;
}

static int launch_dwconv_adaln_ws(const float *x, const float *dw_w, const float *dw_b, const float *scale,
                                  const float *shift, float *y, float *ws, int B, int C, int L, float eps, cudaStream_t st) {
    LINA_REQUIRE(x && scale && shift && y && ws, LINA_ERR_BAD_ARG, "dwconv_adaln: null pointer");
    LINA_REQUIRE(B > 0 && C > 0 && L > 0, LINA_ERR_BAD_ARG, "dwconv_adaln: bad size");
    const int nblk = (C + DT_C - 1) / DT_C;
    LINA_REQUIRE(B <= 65535 && nblk <= 65535, LINA_ERR_UNSUPPORTED, "dwconv_adaln: grid too large");
    LINA_REQUIRE((uintptr_t)scale % 16 == 0 && (uintptr_t)shift % 16 == 0, LINA_ERR_UNSUPPORTED, "dwconv_adaln: scale / shift alignment");
    dim3 grid((L + DT_L - 1) / DT_L, nblk, B);
    if (dw_w != nullptr) dwconv_t_stats_kernel<true><<<grid, 256, 0, st>>>(x, dw_w, dw_b, y, ws, C, L, nblk);
    else dwconv_t_stats_kernel<false><<<grid, 256, 0, st>>>(x, dw_w, dw_b, y, ws, C, L, nblk);
    LINA_LAUNCH_OK("dwconv_t_stats_kernel");
    const long long rows = (long long)B * L;
    adaln_apply_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(y, ws, scale, shift, rows, C, nblk, eps);
    LINA_LAUNCH_OK("adaln_apply_kernel");
    return LINA_OK;
}

extern "C" size_t lina_codec_dwconv_adaln_workspace_bytes(int B, int C, int L) {
    return (size_t)B * L * ((C + DT_C - 1) / DT_C) * 2 * sizeof(float);
}

extern "C" int lina_codec_dwconv_adaln_ws(const float *x, const float *dw_w, const float *dw_b, const float *scale,
                                          const float *shift, float *y, float *ws, int B, int C, int L, float eps,
                                          void *stream) {
    return launch_dwconv_adaln_ws(x, dw_w, dw_b, scale, shift, y, ws, B, C, L, eps, (cudaStream_t)stream);
}

extern "C" int lina_codec_layernorm_t_ws(const float *x, const float *gamma, const float *beta, float *y, float *ws, int B,
                                         int C, int L, float eps, void *stream) {
    return launch_dwconv_adaln_ws(x, nullptr, nullptr, gamma, beta, y, ws, B, C, L, eps, (cudaStream_t)stream);
}

extern "C" int lina_codec_dwconv_adaln(const float *x, const float *dw_w, const float *dw_b, const float *scale,
                                       const float *shift, float *y, int B, int C, int L, float eps, void *stream) {
    return launch_dwconv_adaln(x, dw_w, dw_b, scale, shift, y, B, C, L, eps, (cudaStream_t)stream);
}

extern "C" int lina_codec_layernorm_t(const float *x, const float *gamma, const float *beta, float *y, int B, int C,
                                      int L, float eps, void *stream) {
    return launch_dwconv_adaln(x, nullptr, nullptr, gamma, beta, y, B, C, L, eps, (cudaStream_t)stream);
}

extern "C" int lina_codec_scale_residual_t(const float *h, const float *gamma, const float *res, float *out, int B,
                                           int C, int L, void *stream) {
    LINA_REQUIRE(h && res && out, LINA_ERR_BAD_ARG, "scale_residual_t: null pointer");
    LINA_REQUIRE(B > 0 && C > 0 && L > 0, LINA_ERR_BAD_ARG, "scale_residual_t: bad size");
    LINA_REQUIRE(B <= 65535 && (C + 31) / 32 <= 65535, LINA_ERR_UNSUPPORTED, "scale_residual_t: grid too large");
    if (C % 4 == 0 && L % 2 == 0 && (uintptr_t)h % 16 == 0 && (uintptr_t)res % 8 == 0 && (uintptr_t)out % 8 == 0) {
        dim3 g64((L + 63) / 64, (C + 63) / 64, B);
        scale_residual_t64_kernel<<<g64, 256, 0, (cudaStream_t)stream>>>(h, gamma, res, out, C, L);
        LINA_LAUNCH_OK("scale_residual_t64_kernel");
        return LINA_OK;
    }
    dim3 grid((L + 31) / 32, (C + 31) / 32, B);
    scale_residual_t_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(h, gamma, res, out, C, L);
    LINA_LAUNCH_OK("scale_residual_t_kernel");
    return LINA_OK;
}

extern "C" size_t lina_codec_istft_workspace_bytes(int B, int L, int n_fft) {
    return (size_t)B * L * n_fft * sizeof(float);
}

extern "C" int lina_codec_istft_head(const float *h, const float *window, float *wav, void *ws, int B, int L,
                                     int n_fft, int hop, void *stream) {
    return lina_codec_istft_head_ld(h, (long long)n_fft + 2, window, wav, ws, B, L, n_fft, hop, stream);
}

extern "C" int lina_codec_istft_head_ld(const float *h, long long ldh, const float *window, float *wav, void *ws, int B, int L,
                                        int n_fft, int hop, void *stream) {
    LINA_REQUIRE(h && window && wav && ws, LINA_ERR_BAD_ARG, "istft_head: null pointer");
    LINA_REQUIRE(ldh >= n_fft + 2, LINA_ERR_BAD_ARG, "istft_head: row stride %lld < n_fft + 2", ldh);
    LINA_REQUIRE(B > 0 && L > 0 && n_fft > 0 && hop > 0, LINA_ERR_BAD_ARG, "istft_head: bad size");
    LINA_REQUIRE(n_fft % 4 == 0 && hop <= n_fft && (n_fft - hop) % 2 == 0, LINA_ERR_UNSUPPORTED,
                 "istft_head: need n_fft %% 4 == 0, hop <= n_fft, (n_fft-hop) even (n_fft=%d hop=%d)", n_fft, hop);
    FftPlan plan;
    LINA_REQUIRE(make_plan(n_fft / 2, plan), LINA_ERR_UNSUPPORTED,
                 "istft_head: n_fft/2=%d must factor into 2,3,4,5", n_fft / 2);
    const int M = n_fft / 2;
    const size_t smem = ((size_t)4 * M + 1) * 2 * sizeof(float);
    LINA_REQUIRE(smem <= 48 * 1024, LINA_ERR_UNSUPPORTED, "istft_head: n_fft=%d too large", n_fft);
    LINA_REQUIRE(B <= 65535, LINA_ERR_UNSUPPORTED, "istft_head: B > 65535");
    cudaStream_t st = (cudaStream_t)stream;
    const int nframes = B * L;
    if (g_lina_variant[8] != 2 && n_fft == fft640::N && (uintptr_t)ws % 8 == 0) {     // key 8 = 2: the generic FFT (A/B)
        static thread_local uint64_t configured = 0;
        if (lina_first_use_on_device(&configured))
            LINA_CUDA_OK(cudaFuncSetAttribute(istft_frames_1280_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F640_SMEM));
        const int want = (nframes + F640_WARPS - 1) / F640_WARPS;
        istft_frames_1280_kernel<<<want < 148 * 2 ? want : 148 * 2, F640_WARPS * 32, F640_SMEM, st>>>(h, ldh, window, (float *)ws, nframes);
        LINA_LAUNCH_OK("istft_frames_1280_kernel");
        dim3 g2((L * hop + 255) / 256, B);
        istft_ola_kernel<<<g2, 256, 0, st>>>((const float *)ws, window, wav, L, n_fft, hop);
        LINA_LAUNCH_OK("istft_ola_kernel");
        return LINA_OK;
    }
    const int grid = nframes < 148 * 16 ? nframes : 148 * 16;
    istft_frames_kernel<<<grid, 128, smem, st>>>(h, ldh, window, (float *)ws, nframes, plan);
    LINA_LAUNCH_OK("istft_frames_kernel");
    dim3 g2((L * hop + 255) / 256, B);
    istft_ola_kernel<<<g2, 256, 0, st>>>((const float *)ws, window, wav, L, n_fft, hop);
    LINA_LAUNCH_OK("istft_ola_kernel");
    return LINA_OK;
}
