// Weight-streaming linear layers for the autoregressive step (M = batch <= 32 rows): y = act(LN?(x) W^T + b).
//
// One decode step of the d1024 model runs 52 linears over 13 blocks (model/gla.py:91-99,225, model/base_blocks.py:45-50) with
// M = 32 rows: 0.4 GFLOP against 328 MB of weights -- pure weight bandwidth, and (measured) launch-latency bound: 155 launches
// of ~5.7 us each.  This kernel does what a library GEMM cannot: the step's row-wise stage in FRONT of the linear (residual
// add + LayerNorm, MixingBlock :65-69) runs as its prologue and the stage BEHIND it (SwiGLU's silu(gate) * u, :47-49) as its
// epilogue, so a block is 6 launches instead of 10, and every SM streams its slice of the weight exactly once:
//
//   grid  = ceil(N / NT) CTAs, NT = 8 * NT8 output columns each (N = 6160 -> 129 CTAs of 48 columns, N = 1024 -> 128 of 8);
//   A     = the (normalised) 32 x K activation block, bf16, whole in shared memory (every CTA recomputes the prologue: 64 KB of
//           L2 reads and a few warp reductions -- cheaper than a launch);
//   K     is split over the 8 warps in 32-wide blocks; per block a lane loads ONE 16-byte chunk per weight row (8 rows x 64 B
//           per n8 tile: whole sectors) and feeds mma.sync.m16n8k16 (bf16, fp32 accumulate) through a k-permutation that makes
//           the 16-byte chunk the fragment -- no ldmatrix, no shared-memory staging of weights;
//   reduce the 8 partial tiles through shared memory, add the bias, apply the epilogue, store bf16.
// HBM-bound by construction; tensor cores are mma.sync here on purpose: the work is 0.1 % of a tcgen05 tile's worth.
#include "common.cuh"

namespace {

constexpr int SK_WARPS = 8, SK_THREADS = SK_WARPS * 32, SK_M = 32;

struct SkinnyArgs {
    const bf16 *x;          // [M, ldx]
    const bf16 *delta;      // [M, ldd] added to x before the LayerNorm (nullable)
    const bf16 *ln_g, *ln_b;
    bf16 *sum_out;          // x + delta, [M, K] contiguous (written by CTA 0 when delta != nullptr)
    const bf16 *W;          // [N (+ second half at row `pair_off`), ldw]
    const bf16 *bias;       // [N] / [2 * pair rows] (nullable)
    bf16 *out;              // [M, ldo]
    long long ldx, ldd, ldw, ldo;
    int M, N, K, pair_off;  // pair_off > 0: SwiGLU -- output column j needs weight rows j (gate) and pair_off + j (u)
    float eps;
};

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float bf16r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// PRO 0: A = x.  PRO 1: A = LayerNorm(x + delta) * g + b (rounded exactly like lina_add_layernorm).
// EPI 0: out = bf16(acc + bias).  EPI 1: out = bf16(silu(bf16(gate + bias)) * bf16(u + bias)) (like linear + lina_swiglu_act).
template <int NT8, int PRO, int EPI>
__global__ void __launch_bounds__(SK_THREADS) skinny_linear_kernel(const SkinnyArgs a) {
    constexpr int MAXB = NT8 >= 3 ? 4 : 8;                          // K blocks of 32 per warp: K <= 256 * MAXB (host-checked)
    extern __shared__ __align__(16) uint8_t smem[];
    const int K = a.K, Kp = (K + 31) / 32 * 32, lda = Kp + 8;       // +8 bf16: rows shift by 4 banks
    bf16 *As = reinterpret_cast<bf16 *>(smem);                      // [SK_M][lda]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    const int n0 = blockIdx.x * (EPI == 1 ? NT8 / 2 : NT8) * 8;      // first OUTPUT column of this CTA
    const int nblk = Kp / 32;

    // ---- phase 0: ALL of this warp's weight chunks in flight (they depend on nothing this kernel computes) ----------------
    // warp w takes the 32-wide K blocks w, w + 8, ...; a lane loads one 16-byte chunk per weight row and block
    uint4 bw[MAXB][NT8];
    {
        const bf16 *wrow[NT8];
#pragma unroll
        for (int t = 0; t < NT8; ++t) {
            int row;
            if (EPI == 1) row = (t < NT8 / 2) ? n0 + t * 8 + g : a.pair_off + n0 + (t - NT8 / 2) * 8 + g;
            else row = n0 + t * 8 + g;
            const int lim = EPI == 1 ? ((t < NT8 / 2) ? a.N : a.pair_off + a.N) : a.N;
            if (row >= lim) row = lim - 1;                          // clamp: the columns are masked at the store
            wrow[t] = a.W + (size_t)row * a.ldw + q * 8;
        }
#pragma unroll
        for (int i = 0; i < MAXB; ++i) {
            const int kb = warp + i * SK_WARPS;
#pragma unroll
            for (int t = 0; t < NT8; ++t)
                bw[i][t] = (kb < nblk && kb * 32 + q * 8 < K) ? __ldg(reinterpret_cast<const uint4 *>(wrow[t] + kb * 32))
                                                              : make_uint4(0u, 0u, 0u, 0u);
        }
    }

    // ---- phase 1: the activation block into shared memory ------------------------------------------------------------------
    if (PRO == 0) {
        // plain copy, 8 chunks in flight per thread
        const int cpr = Kp / 8;                                      // 16-byte chunks per row
        for (int i0 = tid; i0 < SK_M * cpr; i0 += 8 * SK_THREADS) {
            uint4 xv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * SK_THREADS, r = i / cpr, c = (i - r * cpr) * 8;
                xv[u] = make_uint4(0u, 0u, 0u, 0u);
                if (i < SK_M * cpr && r < a.M && c < K) xv[u] = *reinterpret_cast<const uint4 *>(a.x + (size_t)r * a.ldx + c);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * SK_THREADS, r = i / cpr, c = (i - r * cpr) * 8;
                if (i < SK_M * cpr) *reinterpret_cast<uint4 *>(As + (size_t)r * lda + c) = xv[u];
            }
        }
    } else {
        // add + LayerNorm, a warp per row, TWO rows at a time with every load of both rows issued before the first use
        // (K <= 1024: 4 chunks of 8 per lane and row); the LayerNorm's weight / bias chunks are loaded once per lane
        uint4 gv[4], bv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = (j * 32 + lane) * 8;
            gv[j] = bv[j] = make_uint4(0u, 0u, 0u, 0u);
            if (c < K) { gv[j] = __ldg(reinterpret_cast<const uint4 *>(a.ln_g + c)); bv[j] = __ldg(reinterpret_cast<const uint4 *>(a.ln_b + c)); }
        }
#pragma unroll 1
        for (int r0 = warp; r0 < SK_M; r0 += 2 * SK_WARPS) {
            uint4 xv[2][4], dv[2][4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = r0 + h * SK_WARPS;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = (j * 32 + lane) * 8;
                    xv[h][j] = dv[h][j] = make_uint4(0u, 0u, 0u, 0u);
                    if (r < a.M && c < K) {
                        xv[h][j] = *reinterpret_cast<const uint4 *>(a.x + (size_t)r * a.ldx + c);
                        if (a.delta != nullptr) dv[h][j] = *reinterpret_cast<const uint4 *>(a.delta + (size_t)r * a.ldd + c);
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = r0 + h * SK_WARPS;
                bf16 *dst = As + (size_t)r * lda;
                float v[4][8];
                float sm = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = (j * 32 + lane) * 8;
                    const bf16 *xe = reinterpret_cast<const bf16 *>(&xv[h][j]), *de = reinterpret_cast<const bf16 *>(&dv[h][j]);
                    if (a.delta != nullptr) {
                        uint4 sv;
                        bf16 *se = reinterpret_cast<bf16 *>(&sv);
#pragma unroll
                        for (int i = 0; i < 8; ++i) { se[i] = __float2bfloat16_rn(__bfloat162float(de[i]) + __bfloat162float(xe[i])); v[j][i] = __bfloat162float(se[i]); }
                        if (blockIdx.x == 0 && r < a.M && c < K) *reinterpret_cast<uint4 *>(a.sum_out + (size_t)r * K + c) = sv;
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[j][i] = __bfloat162float(xe[i]);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) sm += v[j][i];              // chunks past K hold zeros
                }
                const float mean = warp_sum(sm) / (float)K;
                float ss = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((j * 32 + lane) * 8 < K) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { const float d = v[j][i] - mean; ss = fmaf(d, d, ss); }
                    }
                const float rstd = rsqrtf(warp_sum(ss) / (float)K + a.eps);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = (j * 32 + lane) * 8;
                    if (c >= Kp) continue;
                    uint4 ov = make_uint4(0u, 0u, 0u, 0u);
                    if (r < a.M && c < K) {
                        const bf16 *ge = reinterpret_cast<const bf16 *>(&gv[j]), *be = reinterpret_cast<const bf16 *>(&bv[j]);
                        bf16 *oe = reinterpret_cast<bf16 *>(&ov);
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            oe[i] = __float2bfloat16_rn((v[j][i] - mean) * rstd * __bfloat162float(ge[i]) + __bfloat162float(be[i]));
                    }
                    *reinterpret_cast<uint4 *>(dst + c) = ov;
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 3: MMAs ------------------------------------------------------------------------------------------------------
    float acc[2][NT8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int t = 0; t < NT8; ++t)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][t][i] = 0.f;
#pragma unroll
    for (int i = 0; i < MAXB; ++i) {
        const int kb = warp + i * SK_WARPS;
        if (kb < nblk) {
            const int kc = kb * 32 + q * 8;
            uint4 af[2][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                af[mt][0] = *reinterpret_cast<const uint4 *>(As + (size_t)(mt * 16 + g) * lda + kc);
                af[mt][1] = *reinterpret_cast<const uint4 *>(As + (size_t)(mt * 16 + g + 8) * lda + kc);
            }
            // k permutation: physical elements 0..3 of every lane's chunk are MMA step 0 (k pairs {2q,2q+1} and {2q+8,2q+9}),
            // elements 4..7 step 1; A and W use the same chunks, so the dot product covers each physical k exactly once
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int t = 0; t < NT8; ++t) {
                    mma16816(acc[mt][t], af[mt][0].x, af[mt][1].x, af[mt][0].y, af[mt][1].y, bw[i][t].x, bw[i][t].y);
                    mma16816(acc[mt][t], af[mt][0].z, af[mt][1].z, af[mt][0].w, af[mt][1].w, bw[i][t].z, bw[i][t].w);
                }
        }
    }
    __syncthreads();                                                // A is dead: its shared memory becomes the reduction buffer

    // ---- cross-warp reduction + epilogue --------------------------------------------------------------------------------
    constexpr int NT = NT8 * 8;
    float *red = reinterpret_cast<float *>(smem);                   // [SK_WARPS][SK_M][NT]
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int t = 0; t < NT8; ++t) {
            float *r0 = red + ((size_t)warp * SK_M + mt * 16 + g) * NT + t * 8 + 2 * q;
            r0[0] = acc[mt][t][0]; r0[1] = acc[mt][t][1];
            r0[8 * NT] = acc[mt][t][2]; r0[8 * NT + 1] = acc[mt][t][3];
        }
    __syncthreads();
    constexpr int NOUT = EPI == 1 ? NT / 2 : NT;
    for (int i = tid; i < SK_M * NOUT; i += SK_THREADS) {
        const int m = i / NOUT, c = i - m * NOUT, n = n0 + c;
        if (m >= a.M || n >= a.N) continue;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int w = 0; w < SK_WARPS; ++w) {
            s0 += red[((size_t)w * SK_M + m) * NT + c];
            if (EPI == 1) s1 += red[((size_t)w * SK_M + m) * NT + NT / 2 + c];
        }
        if (EPI == 0) {
            if (a.bias != nullptr) s0 += __bfloat162float(a.bias[n]);
            a.out[(size_t)m * a.ldo + n] = __float2bfloat16_rn(s0);
        } else {
            if (a.bias != nullptr) { s0 += __bfloat162float(a.bias[n]); s1 += __bfloat162float(a.bias[a.pair_off + n]); }
            const float gt = bf16r(s0), u = bf16r(s1);
            a.out[(size_t)m * a.ldo + n] = __float2bfloat16_rn(gt / (1.f + __expf(-gt)) * u);
        }
    }
}

template <int NT8, int PRO, int EPI>
int launch_skinny(const SkinnyArgs &a, cudaStream_t st) {
    const int Kp = (a.K + 31) / 32 * 32;
    const size_t smem_a = (size_t)SK_M * (Kp + 8) * sizeof(bf16), smem_r = (size_t)SK_WARPS * SK_M * NT8 * 8 * sizeof(float);
    const size_t smem = smem_a > smem_r ? smem_a : smem_r;
    static thread_local uint64_t configured = 0;
    if (lina_first_use_on_device(&configured))
        LINA_CUDA_OK(cudaFuncSetAttribute(skinny_linear_kernel<NT8, PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    LINA_REQUIRE(smem <= 200 * 1024 && Kp <= 256 * (NT8 >= 3 ? 4 : 8) && (PRO == 0 || a.K <= 1024), LINA_ERR_UNSUPPORTED,
                 "skinny_linear: K = %d too large for %d columns per CTA", a.K, NT8 * 8);
    const int cols = (EPI == 1 ? NT8 / 2 : NT8) * 8;
    const int grid = (a.N + cols - 1) / cols;
    skinny_linear_kernel<NT8, PRO, EPI><<<grid, SK_THREADS, smem, st>>>(a);
    LINA_LAUNCH_OK("skinny_linear_kernel");
    return LINA_OK;
}

bool al16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

}  // namespace

extern "C" int lina_skinny_linear_max_rows(void) { return SK_M; }

extern "C" int lina_skinny_linear(const void *x, long long ldx, const void *delta, long long ldd, const void *ln_gamma,
                                  const void *ln_beta, float ln_eps, void *sum_out, const void *W, long long ldw, const void *bias, void *out,
                                  long long ldo, int M, int N, int K, int swiglu_pair_offset, void *stream) {
    LINA_REQUIRE(x && W && out, LINA_ERR_BAD_ARG, "skinny_linear: null pointer");
    LINA_REQUIRE(M > 0 && M <= SK_M && N > 0 && K > 0, LINA_ERR_UNSUPPORTED, "skinny_linear: needs 1 <= M <= %d rows (M %d)", SK_M, M);
    LINA_REQUIRE(K % 8 == 0 && K <= 2048 && ldx % 8 == 0 && ldw % 8 == 0 && ldx >= K && ldw >= K && ldo >= N, LINA_ERR_UNSUPPORTED,
                 "skinny_linear: K %% 8 == 0, K <= 2048, row strides multiples of 8 covering a row (K %d ldx %lld ldw %lld)", K, ldx, ldw);
    LINA_REQUIRE(al16(x) && al16(W) && (delta == nullptr || al16(delta)) && (sum_out == nullptr || al16(sum_out)) &&
                     (ln_gamma == nullptr || al16(ln_gamma)) && (ln_beta == nullptr || al16(ln_beta)),
                 LINA_ERR_UNSUPPORTED, "skinny_linear: tensors must be 16-byte aligned");
    LINA_REQUIRE((ln_gamma == nullptr) == (ln_beta == nullptr), LINA_ERR_BAD_ARG, "skinny_linear: LayerNorm needs weight and bias");
    LINA_REQUIRE(delta == nullptr || (ln_gamma != nullptr && sum_out != nullptr && ldd % 8 == 0 && ldd >= K), LINA_ERR_BAD_ARG,
                 "skinny_linear: the residual add is part of the LayerNorm prologue and needs sum_out (and a row stride %% 8)");
    LINA_REQUIRE(swiglu_pair_offset >= 0 && (swiglu_pair_offset == 0 || swiglu_pair_offset >= N), LINA_ERR_BAD_ARG,
                 "skinny_linear: swiglu_pair_offset must be 0 or >= N");
    SkinnyArgs a{};
    a.x = (const bf16 *)x; a.delta = (const bf16 *)delta; a.ln_g = (const bf16 *)ln_gamma; a.ln_b = (const bf16 *)ln_beta;
    a.sum_out = (bf16 *)sum_out; a.W = (const bf16 *)W; a.bias = (const bf16 *)bias; a.out = (bf16 *)out;
    a.ldx = ldx; a.ldd = ldd; a.ldw = ldw; a.ldo = ldo; a.M = M; a.N = N; a.K = K; a.pair_off = swiglu_pair_offset; a.eps = ln_eps;
    cudaStream_t st = (cudaStream_t)stream;
    const bool ln = ln_gamma != nullptr;
    if (swiglu_pair_offset > 0)                 // 8 output columns per CTA: one gate tile + one u tile
        return ln ? launch_skinny<2, 1, 1>(a, st) : launch_skinny<2, 0, 1>(a, st);
    // columns per CTA so that the grid is about one CTA per SM
    if (N >= 4096 && K <= 1024) return ln ? launch_skinny<6, 1, 0>(a, st) : launch_skinny<6, 0, 0>(a, st);
    if (N >= 2048) return ln ? launch_skinny<2, 1, 0>(a, st) : launch_skinny<2, 0, 0>(a, st);
    return ln ? launch_skinny<1, 1, 0>(a, st) : launch_skinny<1, 0, 0>(a, st);
}
