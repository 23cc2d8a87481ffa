// WavTokenizer decode, channels-last edition: the streaming stages BETWEEN the tensor-core contractions of csrc/gemm_sm100.cu.
//
// Every activation lives as [B, L, C] (channels contiguous), which is (a) the K-major A operand the GEMM's TMA wants -- so the
// reference's transposes around every ConvNeXt block / norm (DEC/modules.py:50,58, DEC/models.py:231-233) disappear -- and
// (b) one fully coalesced row per warp for the row-wise stages.  A stage that feeds a contraction writes the bf16 SPLIT of
// its fp32 result (hi [, mid], lo: see gemm_sm100.cu) instead of the fp32 tensor: same bytes, no extra pass.
//
//   lina_codec_cl_gather      codes -> codebook rows summed over quantizers -> split parts   (DEC/pretrained.py:231-237)
//   lina_codec_cl_gn_partials per-(batch, 32-row tile, group) Welford partials of GroupNorm(32, C)   (DEC/models.py:15-16)
//   lina_codec_cl_rows        [depthwise conv k=7] -> [GroupNorm apply] -> [swish] -> [LayerNorm * scale + shift] per row,
//                             fp32 and / or split out   (ConvNeXtBlock :48-51, AdaLayerNorm :81-86, ResnetBlock :61-70,
//                             AttnBlock :109, final_layer_norm, pos_net[5] + backbone.norm)
//   lina_codec_cl_softmax     row softmax of the attention scores -> split parts   (DEC/models.py:119-120)
#include "common.cuh"

namespace {

constexpr int GN_ROWS = 32;          // rows per partial tile
constexpr int MAXP = 3;

__device__ __forceinline__ void split_store(float v, bf16 *const *parts, int nparts, size_t idx) {
#pragma unroll
    for (int p = 0; p < MAXP; ++p) {
        if (p >= nparts) break;
        const bf16 h = __float2bfloat16_rn(v);
        parts[p][idx] = h;
        v -= __bfloat162float(h);
    }
}

struct Parts { bf16 *p[MAXP]; int n; };

// 4 consecutive channels -> 8-byte stores per part
__device__ __forceinline__ void split_store4(float4 v, const Parts &o, size_t idx) {
    float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int p = 0; p < MAXP; ++p) {
        if (p >= o.n) break;
        uint32_t w0, w1;                                 // one cvt per pair; the remainders feed the next part
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w0) : "f"(a[1]), "f"(a[0]));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w1) : "f"(a[3]), "f"(a[2]));
        a[0] -= __uint_as_float(w0 << 16); a[1] -= __uint_as_float(w0 & 0xFFFF0000u);
        a[2] -= __uint_as_float(w1 << 16); a[3] -= __uint_as_float(w1 & 0xFFFF0000u);
        *reinterpret_cast<uint2 *>(o.p[p] + idx) = make_uint2(w0, w1);
    }
}

// ---- codes -> features ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cl_gather_kernel(const int64_t *__restrict__ codes, const float *__restrict__ books, Parts out, float *__restrict__ out_f32,
                 int Kq, int B, int L, int bins, int C) {
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B * L) return;
    for (int c = lane * 4; c < C; c += 128) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < Kq; ++q) {
            long long code = codes[(size_t)q * B * L + warp];
            code = code < 0 ? 0 : (code >= bins ? bins - 1 : code);
            const float4 t = *reinterpret_cast<const float4 *>(books + ((size_t)q * bins + code) * C + c);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        const size_t idx = (size_t)warp * C + c;
        if (out_f32 != nullptr) *reinterpret_cast<float4 *>(out_f32 + idx) = acc;
        split_store4(acc, out, idx);
    }
}

// ---- GroupNorm statistics ----------------------------------------------------------------------------------------------
// grid (tiles, B), C threads: thread c runs Welford over its channel's GN_ROWS rows, then thread g merges the channels of
// group g (Chan) and writes (count, mean, M2) to partials[b][tile][g].
__global__ void __launch_bounds__(1024)
cl_gn_partials_kernel(const float *__restrict__ x, float *__restrict__ partials, int L, int C, int G) {
    extern __shared__ float sm[];                    // [C][2] mean, M2
    const int b = blockIdx.y, tile = blockIdx.x, c = threadIdx.x;
    const int l0 = tile * GN_ROWS, nrow = min(GN_ROWS, L - l0);
    const float *xp = x + ((size_t)b * L + l0) * C + c;
    // shifted sums (shift = the channel's first sample of the tile): sum (x - K), sum (x - K)^2 carry no cancellation
    // problem in fp32 and need no division per element; (mean, M2) follow exactly as for Welford
    const float K0 = xp[0];
    float s1 = 0.f, s2 = 0.f;
    if (nrow == GN_ROWS) {                      // full tile: every load of the tile in flight at once
        float d[GN_ROWS];
#pragma unroll
        for (int r = 0; r < GN_ROWS; ++r) d[r] = xp[(size_t)r * C];
#pragma unroll
        for (int r = 0; r < GN_ROWS; ++r) { const float t = d[r] - K0; s1 += t; s2 = fmaf(t, t, s2); }
    } else {
        for (int r = 0; r < nrow; ++r) {
            const float t = xp[(size_t)r * C] - K0;
            s1 += t;
            s2 = fmaf(t, t, s2);
        }
    }
    const float mean = K0 + s1 / (float)nrow;
    const float m2 = fmaxf(s2 - s1 * s1 / (float)nrow, 0.f);
    sm[2 * c] = mean; sm[2 * c + 1] = m2;
    __syncthreads();
    if (c < G) {
        const int cpg = C / G;
        float n = 0.f, mu = 0.f, M2 = 0.f;
        for (int i = 0; i < cpg; ++i) {
            const float nb = (float)nrow, mb = sm[2 * (c * cpg + i)], Mb = sm[2 * (c * cpg + i) + 1];
            const float nn = n + nb, d = mb - mu;
            mu += d * nb / nn;
            M2 += Mb + d * d * n * nb / nn;
            n = nn;
        }
        float *o = partials + (((size_t)b * gridDim.x + tile) * G + c) * 3;
        o[0] = n; o[1] = mu; o[2] = M2;
    }
}

// ---- the row kernel ----------------------------------------------------------------------------------------------------
struct RowArgs {
    const float *x;
    const float *dw_w, *dw_b;               // [C][7], [C]   (nullptr: no depthwise conv)
    const float *gn_partials, *gn_w, *gn_b; // partials [B][tiles][G][3], affine [C]   (nullptr: no GroupNorm)
    const float *ln_scale, *ln_shift;       // [C]   (nullptr: no LayerNorm)
    float *out_f32;
    Parts out;
    int B, L, C, G, gn_tiles, swish;
    float gn_eps, ln_eps;
};

constexpr int ROW_WARPS = 8;
constexpr int ROWS_PER_WARP = 4;                 // consecutive rows of one warp (the CTA covers 32 consecutive rows)

template <int VEC, int MAXJ>      // VEC 4: C % 128 == 0 (float4 per lane and step), 1: any C; MAXJ: steps per lane, C <= 32 * VEC * MAXJ
__global__ void __launch_bounds__(ROW_WARPS * 32, (VEC * MAXJ <= 24) ? 3 : 2)
cl_rows_kernel(const RowArgs a) {
    __shared__ float gstat[64][2];                              // GroupNorm (mean, rstd) of this CTA's batch (G <= 64)
    extern __shared__ float dyn[];
    // dynamic shared memory: [gnA | gnB] (2C, GroupNorm folded to x * A[c] + B[c] for this CTA's batch) | [lnS | lnT] (2C) |
    // depthwise taps transposed [7][C] + bias [C]; absent stages take no space (offsets from the host-side layout)
    float *gnA = dyn, *lnS = dyn + (a.gn_partials != nullptr ? 2 * a.C : 0);
    float *dws = lnS + (a.ln_scale != nullptr ? 2 * a.C : 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = a.C, L = a.L;
    constexpr int CTA_ROWS = ROW_WARPS * ROWS_PER_WARP;
    const int tiles_per_b = (L + CTA_ROWS - 1) / CTA_ROWS;      // CTA = CTA_ROWS consecutive rows of ONE batch
    const int b = blockIdx.x / tiles_per_b;
    const int lbase = (blockIdx.x - b * tiles_per_b) * CTA_ROWS + warp * ROWS_PER_WARP;
    if (a.dw_w != nullptr) {
        for (int i = threadIdx.x; i < 7 * C; i += ROW_WARPS * 32) { const int t = i / C, c = i - t * C; dws[i] = a.dw_w[c * 7 + t]; }
        for (int i = threadIdx.x; i < C; i += ROW_WARPS * 32) dws[7 * C + i] = a.dw_b != nullptr ? a.dw_b[i] : 0.f;
    }
    if (a.gn_partials != nullptr) {
        if (threadIdx.x < a.G) {
            float n = 0.f, mu = 0.f, M2 = 0.f;
            for (int t = 0; t < a.gn_tiles; ++t) {
                const float *pp = a.gn_partials + (((size_t)b * a.gn_tiles + t) * a.G + threadIdx.x) * 3;
                const float nb = pp[0], mb = pp[1], Mb = pp[2];
                const float nn = n + nb, d = mb - mu;
                mu += d * nb / nn;
                M2 += Mb + d * d * n * nb / nn;
                n = nn;
            }
            gstat[threadIdx.x][0] = mu;
            gstat[threadIdx.x][1] = rsqrtf(M2 / n + a.gn_eps);
        }
        __syncthreads();
        const int cpg = C / a.G;
        for (int c = threadIdx.x; c < C; c += ROW_WARPS * 32) {
            const int g = c / cpg;
            const float A = gstat[g][1] * a.gn_w[c];
            gnA[c] = A;
            gnA[C + c] = a.gn_b[c] - gstat[g][0] * A;
        }
    }
    if (a.ln_scale != nullptr) {
        for (int c = threadIdx.x; c < C; c += ROW_WARPS * 32) { lnS[c] = a.ln_scale[c]; lnS[C + c] = a.ln_shift != nullptr ? a.ln_shift[c] : 0.f; }
    }
    __syncthreads();
    const int nj = C / (32 * VEC);
#pragma unroll 1
    for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
    const int l = lbase + rr;
    if (l >= L) break;
    const float *xr = a.x + ((size_t)b * L + l) * C;
    float v[MAXJ * VEC];
    // load (+ depthwise conv over the 7 neighbouring rows; rows outside [0, L) are the conv's zero padding)
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
        if (j >= nj) break;
        const int c = (j * 32 + lane) * VEC;
        if (a.dw_w == nullptr) {
            if (VEC == 4) {
                const float4 t = *reinterpret_cast<const float4 *>(xr + c);
                v[j * VEC] = t.x; v[j * VEC + 1 % VEC] = t.y; v[j * VEC + 2 % VEC] = t.z; v[j * VEC + 3 % VEC] = t.w;
            } else {
                v[j * VEC] = xr[c];
            }
        } else {
            float acc[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] = dws[7 * C + c + i];
#pragma unroll
            for (int t = 0; t < 7; ++t) {
                const int ll = l + t - 3;
                if (ll < 0 || ll >= L) continue;
                const float *xp = a.x + ((size_t)b * L + ll) * C + c;
                if (VEC == 4) {
                    const float4 xv = *reinterpret_cast<const float4 *>(xp);
                    const float4 wv = *reinterpret_cast<const float4 *>(dws + t * C + c);
                    acc[0] = fmaf(wv.x, xv.x, acc[0]);
                    acc[1 % VEC] = fmaf(wv.y, xv.y, acc[1 % VEC]);
                    acc[2 % VEC] = fmaf(wv.z, xv.z, acc[2 % VEC]);
                    acc[3 % VEC] = fmaf(wv.w, xv.w, acc[3 % VEC]);
                } else {
                    acc[0] = fmaf(dws[t * C + c], xp[0], acc[0]);
                }
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) v[j * VEC + i] = acc[i];
        }
    }
    if (a.gn_partials != nullptr) {
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
            if (j >= nj) break;
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const int c = (j * 32 + lane) * VEC + i;
                v[j * VEC + i] = fmaf(v[j * VEC + i], gnA[c], gnA[C + c]);
            }
        }
    }
    if (a.swish) {                                     // x * sigmoid(x): ex2 + rcp on the MUFU (relative error ~1e-6)
#pragma unroll
        for (int j = 0; j < MAXJ * VEC; ++j) {
            if (j >= nj * VEC) break;
            v[j] = __fdividef(v[j], 1.f + __expf(-v[j]));
        }
    }
    if (a.ln_scale != nullptr) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < MAXJ * VEC; ++j) { if (j >= nj * VEC) break; s += v[j]; }
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < MAXJ * VEC; ++j) { if (j >= nj * VEC) break; const float d = v[j] - mean; q = fmaf(d, d, q); }
        const float rstd = rsqrtf(warp_sum(q) / (float)C + a.ln_eps);
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
            if (j >= nj) break;
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const int c = (j * 32 + lane) * VEC + i;
                v[j * VEC + i] = fmaf((v[j * VEC + i] - mean) * rstd, lnS[c], lnS[C + c]);
            }
        }
    }
    const size_t base = ((size_t)b * L + l) * C;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
        if (j >= nj) break;
        const int c = (j * 32 + lane) * VEC;
        if (VEC == 4) {
            const float4 t = make_float4(v[j * VEC], v[j * VEC + 1 % VEC], v[j * VEC + 2 % VEC], v[j * VEC + 3 % VEC]);
            if (a.out_f32 != nullptr) *reinterpret_cast<float4 *>(a.out_f32 + base + c) = t;
            split_store4(t, a.out, base + c);
        } else {
            if (a.out_f32 != nullptr) a.out_f32[base + c] = v[j * VEC];
            split_store(v[j * VEC], a.out.p, a.out.n, base + c);
        }
    }
    }   // rows of this warp
}

// ---- softmax over the key axis -----------------------------------------------------------------------------------------
// S [rows][ldS] fp32 (row length n) -> P parts [rows][ldP] bf16, columns n..ldP-1 zero.  One warp per row.
__global__ void __launch_bounds__(256)
cl_softmax_kernel(const float *__restrict__ S, Parts out, long long rows, int n, long long ldS, long long ldP) {
    const long long row = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float *s = S + row * ldS;
    float m = -INFINITY;
    for (int j = lane; j < n; j += 32) m = fmaxf(m, s[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int j = lane; j < n; j += 32) sum += expf(s[j] - m);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < (int)ldP; j += 32) {
        const float pv = j < n ? expf(s[j] - m) * inv : 0.f;
        split_store(pv, out.p, out.n, (size_t)(row * ldP + j));
    }
}

bool al16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

int fill_parts(Parts &o, void *const *parts, int nparts) {
    LINA_REQUIRE(nparts >= 0 && nparts <= MAXP, LINA_ERR_BAD_ARG, "codec_cl: 0..3 output parts");
    o.n = nparts;
    for (int i = 0; i < MAXP; ++i) o.p[i] = i < nparts ? (bf16 *)parts[i] : nullptr;
    for (int i = 0; i < nparts; ++i) LINA_REQUIRE(parts[i] != nullptr, LINA_ERR_BAD_ARG, "codec_cl: output part %d is null", i);
    return LINA_OK;
}

}  // namespace

extern "C" int lina_codec_cl_gather(const int64_t *codes, const float *codebooks, float *out_f32, void *const *out_parts,
                                    int n_parts, int Kq, int B, int L, int bins, int C, void *stream) {
    LINA_REQUIRE(codes && codebooks && (out_f32 || n_parts > 0), LINA_ERR_BAD_ARG, "codec_cl_gather: null pointer");
    LINA_REQUIRE(Kq > 0 && B > 0 && L > 0 && bins > 0 && C > 0 && C % 4 == 0, LINA_ERR_BAD_ARG, "codec_cl_gather: bad size (C %% 4)");
    Parts o;
    int rc = fill_parts(o, out_parts, n_parts);
    if (rc) return rc;
    const long long rows = (long long)B * L;
    const long long nblk = (rows * 32 + 255) / 256;
    LINA_REQUIRE(nblk <= 2147483647LL, LINA_ERR_UNSUPPORTED, "codec_cl_gather: grid too large");
    cl_gather_kernel<<<(unsigned)nblk, 256, 0, (cudaStream_t)stream>>>(codes, codebooks, o, out_f32, Kq, B, L, bins, C);
    LINA_LAUNCH_OK("cl_gather_kernel");
    return LINA_OK;
}

extern "C" size_t lina_codec_cl_gn_partials_bytes(int B, int L, int G) {
    return (size_t)B * ((L + GN_ROWS - 1) / GN_ROWS) * G * 3 * sizeof(float);
}

extern "C" int lina_codec_cl_gn_partials(const float *x, float *partials, int B, int L, int C, int G, void *stream) {
    LINA_REQUIRE(x && partials, LINA_ERR_BAD_ARG, "codec_cl_gn_partials: null pointer");
    LINA_REQUIRE(B > 0 && L > 0 && C > 0 && G > 0 && C % G == 0 && C <= 1024 && G <= 64 && B <= 65535, LINA_ERR_UNSUPPORTED,
                 "codec_cl_gn_partials: need C %% G == 0, C <= 1024, G <= 64 (C %d G %d)", C, G);
    dim3 grid((L + GN_ROWS - 1) / GN_ROWS, B);
    cl_gn_partials_kernel<<<grid, C, (size_t)C * 2 * sizeof(float), (cudaStream_t)stream>>>(x, partials, L, C, G);
    LINA_LAUNCH_OK("cl_gn_partials_kernel");
    return LINA_OK;
}

extern "C" int lina_codec_cl_rows(const float *x, const float *dw_w, const float *dw_b, const float *gn_partials,
                                  const float *gn_w, const float *gn_b, int G, float gn_eps, int swish,
                                  const float *ln_scale, const float *ln_shift, float ln_eps, float *out_f32,
                                  void *const *out_parts, int n_parts, int B, int L, int C, void *stream) {
    LINA_REQUIRE(x && (out_f32 || n_parts > 0), LINA_ERR_BAD_ARG, "codec_cl_rows: null pointer");
    LINA_REQUIRE(B > 0 && L > 0 && C > 0 && C % 32 == 0 && C <= 1024, LINA_ERR_UNSUPPORTED,
                 "codec_cl_rows: C must be a multiple of 32 and <= 1024 (C %d)", C);
    LINA_REQUIRE(gn_partials == nullptr || (gn_w && gn_b && G > 0 && G <= 64 && C % G == 0), LINA_ERR_BAD_ARG,
                 "codec_cl_rows: GroupNorm needs weight, bias, 0 < G <= 64, C %% G == 0");
    RowArgs a{};
    int rc = fill_parts(a.out, out_parts, n_parts);
    if (rc) return rc;
    a.x = x; a.dw_w = dw_w; a.dw_b = dw_b; a.gn_partials = gn_partials; a.gn_w = gn_w; a.gn_b = gn_b;
    a.ln_scale = ln_scale; a.ln_shift = ln_shift; a.out_f32 = out_f32;
    a.B = B; a.L = L; a.C = C; a.G = G; a.gn_tiles = (L + GN_ROWS - 1) / GN_ROWS; a.swish = swish;
    a.gn_eps = gn_eps; a.ln_eps = ln_eps;
    constexpr int CTA_ROWS = ROW_WARPS * ROWS_PER_WARP;
    const long long nblk = (long long)B * ((L + CTA_ROWS - 1) / CTA_ROWS);
    LINA_REQUIRE(nblk <= 2147483647LL, LINA_ERR_UNSUPPORTED, "codec_cl_rows: grid too large");
    const size_t dsm = (size_t)((gn_partials != nullptr ? 2 : 0) + (ln_scale != nullptr ? 2 : 0) + (dw_w != nullptr ? 8 : 0)) * C * sizeof(float);
    LINA_REQUIRE(dsm <= 48 * 1024, LINA_ERR_UNSUPPORTED, "codec_cl_rows: C = %d too large for the coefficient tables", C);
    bool vec = C % 128 == 0 && al16(x) && (out_f32 == nullptr || al16(out_f32));
    for (int i = 0; i < n_parts; ++i) vec = vec && (((uintptr_t)out_parts[i] & 7u) == 0);
    if (vec && C <= 768) cl_rows_kernel<4, 6><<<(unsigned)nblk, ROW_WARPS * 32, dsm, (cudaStream_t)stream>>>(a);
    else if (vec) cl_rows_kernel<4, 8><<<(unsigned)nblk, ROW_WARPS * 32, dsm, (cudaStream_t)stream>>>(a);
    else if (C <= 256) cl_rows_kernel<1, 8><<<(unsigned)nblk, ROW_WARPS * 32, dsm, (cudaStream_t)stream>>>(a);
    else cl_rows_kernel<1, 32><<<(unsigned)nblk, ROW_WARPS * 32, dsm, (cudaStream_t)stream>>>(a);
    LINA_LAUNCH_OK("cl_rows_kernel");
    return LINA_OK;
}

extern "C" int lina_codec_cl_softmax(const float *S, void *const *out_parts, int n_parts, long long rows, int n,
                                     long long ldS, long long ldP, void *stream) {
    LINA_REQUIRE(S && n_parts > 0, LINA_ERR_BAD_ARG, "codec_cl_softmax: null pointer / no output");
    LINA_REQUIRE(rows > 0 && n > 0 && ldS >= n && ldP >= n, LINA_ERR_BAD_ARG, "codec_cl_softmax: bad size");
    Parts o;
    int rc = fill_parts(o, out_parts, n_parts);
    if (rc) return rc;
    const long long nblk = (rows * 32 + 255) / 256;
    LINA_REQUIRE(nblk <= 2147483647LL, LINA_ERR_UNSUPPORTED, "codec_cl_softmax: grid too large");
    cl_softmax_kernel<<<(unsigned)nblk, 256, 0, (cudaStream_t)stream>>>(S, o, rows, n, ldS, ldP);
    LINA_LAUNCH_OK("cl_softmax_kernel");
    return LINA_OK;
}
