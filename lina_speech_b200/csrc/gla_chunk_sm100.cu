// Chunkwise-parallel GLA forward on the sm_100a tensor cores (tcgen05 + TMEM), bf16 I/O.
//
// Function (spec: FLA/fla/ops/gla/naive.py:13-44; replaces the Triton kernel families of
// FLA/fla/ops/gla/chunk_fuse.py:21-95,205-247 + chunk_util.py:5-65 and FLA/fla/ops/gla/chunk.py:17-136 +
// FLA/fla/ops/common/chunk_h.py:15-98):
//     S_t = diag(exp(gk_t)) S_{t-1} + k_t^T v_t ;   o_t = scale * q_t S_t
// in chunk form over C = 64 tokens, with G = inclusive cumsum of gk inside the chunk (fp32):
//     q~_t = scale * q_t * exp(G_t)            k~_s = k_s * exp(-G_s)            (bf16 MMA operands)
//     P    = tril(q~ k~^T)                                                  [C x C]
//     o    = q~ S + P v                                                     [C x V]
//     S'   = diag(exp(G_C)) (S + k~^T v)                                    [K x V]
//
// One CTA owns one (batch, head) and a 128-wide slice of V for the whole sequence; the state slice
// lives in TENSOR MEMORY, transposed, for the life of the CTA:
//     ST [128 lanes = v, K cols]  fp32   master state          (accumulator of the k~^T v MMA)
//     SA [128 lanes = v, K/2 cols] bf16  copy of ST            (A operand, read from TMEM, of the q~ S MMA)
//     OT [128 lanes = v, 64 cols]  fp32  o^T of the chunk      P [128 lanes, 64 cols] fp32 scores (rows 64.. unused)
// All four MMAs of a chunk are M = 128 tcgen05.mma (kind::f16, bf16 x bf16 -> fp32), single-thread issued:
//     (0) P  = [q~;k~] k~^T   (rows 64..127 are a by-product; M=128 costs what M=64 would)
//     (1) OT = SA q~^T        (A from TMEM)        (2) OT += v^T P^T        (3) ST += v^T k~
// Shared-memory operands use the SWIZZLE_NONE core-matrix layout [depth/8][row][8] (see sm100.cuh); the same
// k~ bytes serve as K-major operand of (0) and MN-major operand of (3), the same v bytes as the MN-major A of
// (2) and (3).
//
// Warp roles (576 threads):
//   warps 5,6  loaders: cp.async (16 B, zero-filled past T) of the raw q, k rows straight into the operand tile
//              of their stage (even / odd chunks) and of gk into a side tile; warp 7 loads the v tile;
//   WG0        gate pre-pass IN PLACE on the landed tile: cumsum over the chunk, q -> q~, k -> k~ (two stages);
//   warp 4     MMA issuer (one thread);
//   warps 8-11 output epilogue (OT: TMEM -> global);   warps 16,17 causal mask (P: TMEM -> bf16 smem);
//   warps 12-15 state pass (ST *= exp(G_C), refresh SA, final state).
// Synchronisation is mbarrier-only; global-load latency is hidden by the loaders running a stage ahead.
//
// HBM traffic per CTA = q,k,gk once + its v slice + its o slice; q,k,gk are re-read by the V/128 CTAs of
// the same (b,h) (L2 hits).  Algorithmic bytes per token per head: (3K + 2V) * 2.
#include "common.cuh"
#include "sm100.cuh"

using namespace sm100;

int lina_gla_recurrent_fwd_impl(const void *q, const void *k, const void *v, const void *gk, const void *h0,
                                int h0_dtype, void *o, float *ht, int B, int H, int T, int K, int V, int dtype,
                                float scale, void *stream);

namespace {

constexpr int C = 64;            // chunk length (tokens)
constexpr int BV = 128;          // V slice per CTA
constexpr int NTHREADS = 576;      // 18 warps, see the role table below
constexpr uint32_t GV = (C + 1) * 16;     // v tile   : [BV/8][64 rows s (+1 pad)][8]
constexpr uint32_t GP = C * 16;           // P tile   : [C/8][64 rows t][8]
constexpr uint32_t VT_BYTES = (BV / 8) * GV;
constexpr uint32_t PT_BYTES = (C / 8) * GP;
// TMEM columns
constexpr uint32_t COL_ST = 0, COL_OT = 256, COL_P = 320, COL_SA = 384;

template <int K> struct Cfg {
    static constexpr int KC = K / 8;                       // 16-byte groups along K
    static constexpr uint32_t GQK = (128 + 1) * 16;        // qk tile : [KC][128 rows (q~ 0..63, k~ 64..127) (+1 pad)][8]
    static constexpr uint32_t QK_BYTES = ((KC * GQK + 127) / 128) * 128;
    static constexpr uint32_t GG = (C + 1) * 16;           // gk tile : [KC][64 rows (+1 pad)][8]
    static constexpr uint32_t G_BYTES = ((KC * GG + 127) / 128) * 128;
    static constexpr int NRG = 128 / KC;                   // row groups of the pre-pass
    static constexpr int RPG = C / NRG;                    // rows per group
    static constexpr uint32_t OFF_QK = 0;
    static constexpr uint32_t OFF_G = OFF_QK + 2 * QK_BYTES;
    static constexpr uint32_t OFF_V = OFF_G + 2 * G_BYTES;
    static constexpr uint32_t OFF_P = OFF_V + VT_BYTES;               // single v stage
    static constexpr uint32_t OFF_DVEC = OFF_P + PT_BYTES;            // [4][K] fp32
    static constexpr uint32_t OFF_PART = OFF_DVEC + 4 * K * 4;        // [NRG][K] fp32
    static constexpr uint32_t OFF_BAR = OFF_PART + NRG * K * 4;       // mbarriers
    static constexpr uint32_t SMEM = OFF_BAR + 24 * 8 + 16;
    static_assert(SMEM <= 232448, "shared memory budget");
};

enum { B_QK_FULL0 = 0, B_QK_FULL1, B_QK_EMPTY0, B_QK_EMPTY1, B_P_FULL, B_P_TEMPTY, B_PS_FULL, B_PS_EMPTY,
       B_O_FULL, B_O_EMPTY, B_ST_FULL, B_SA_FULL, B_RAW_FULL0, B_RAW_FULL1, B_G_EMPTY0, B_G_EMPTY1, B_V_FULL,
       B_V_EMPTY, B_COUNT };

// mbarrier wait that traps instead of hanging forever (a protocol bug must not wedge the GPU)
__device__ __forceinline__ void wait_bar(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) { printf("gla_chunk_sm100: mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x); __trap(); }
    }
}

// bring-up timeline: role r, item n, event e -> clock64 of CTA (0,0)   (trace == nullptr in production)
constexpr int TR_EV = 4, TR_MAXN = 64;
#define TRACE(role, n, ev) do { if (trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && (n) < TR_MAXN) \
    trace[((role) * TR_MAXN + (n)) * TR_EV + (ev)] = clock64(); } while (0)

__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void unpack8(const uint4 &raw, float *f) {
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

template <int K>
__global__ void __launch_bounds__(NTHREADS, 1)
gla_chunk_fwd_sm100_kernel(const bf16 *__restrict__ q, const bf16 *__restrict__ k, const bf16 *__restrict__ v,
                           const bf16 *__restrict__ gk, const void *__restrict__ h0, int h0_dtype,
                           bf16 *__restrict__ o, float *__restrict__ ht, int T, int V, float scale,
                           long long *__restrict__ trace) {
    using cfg = Cfg<K>;
    constexpr int KC = cfg::KC, NRG = cfg::NRG, RPG = cfg::RPG;
    constexpr uint32_t GQK = cfg::GQK;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + cfg::OFF_BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + cfg::OFF_BAR + 24 * 8);
    float *dvec = reinterpret_cast<float *>(smem + cfg::OFF_DVEC);
    float *part = reinterpret_cast<float *>(smem + cfg::OFF_PART);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bh = blockIdx.y, v0 = blockIdx.x * BV;
    const int n_items = (T + C - 1) / C;
    const size_t qbase = (size_t)bh * T * K, vbase = (size_t)bh * T * V;

    if (tid == 0) {
        mbar_init(&bars[B_QK_FULL0], 128); mbar_init(&bars[B_QK_FULL1], 128);
        mbar_init(&bars[B_QK_EMPTY0], 1); mbar_init(&bars[B_QK_EMPTY1], 1);
        mbar_init(&bars[B_P_FULL], 1); mbar_init(&bars[B_P_TEMPTY], 64);
        mbar_init(&bars[B_PS_FULL], 64); mbar_init(&bars[B_PS_EMPTY], 1);
        mbar_init(&bars[B_O_FULL], 1); mbar_init(&bars[B_O_EMPTY], 128);
        mbar_init(&bars[B_ST_FULL], 1); mbar_init(&bars[B_SA_FULL], 128);
        mbar_init(&bars[B_RAW_FULL0], 32); mbar_init(&bars[B_RAW_FULL1], 32);
        mbar_init(&bars[B_G_EMPTY0], 128); mbar_init(&bars[B_G_EMPTY1], 128);
        mbar_init(&bars[B_V_FULL], 32); mbar_init(&bars[B_V_EMPTY], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 4) {
        // ====================== WG0: gate pre-pass, in place on the landed q / k rows ======================
        const int p = tid;                         // 0..127
        const int c = p % KC, rg = p / KC;
        constexpr uint32_t GG = cfg::GG;
        for (int n = 0; n < n_items; ++n) {
            const int s = n & 1;
            uint8_t *qk_tile = smem + cfg::OFF_QK + s * cfg::QK_BYTES;
            const uint8_t *g_tile = smem + cfg::OFF_G + s * cfg::G_BYTES;
            wait_bar(&bars[B_RAW_FULL0 + s], (n >> 1) & 1);
            if (p == 0) TRACE(0, n, 0);
            // pass A: column sums of gk over this thread's rows
            float csum[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) csum[j] = 0.f;
#pragma unroll
            for (int i = 0; i < RPG; ++i) {
                float g8[8];
                unpack8(*reinterpret_cast<const uint4 *>(g_tile + c * GG + (rg * RPG + i) * 16), g8);
#pragma unroll
                for (int j = 0; j < 8; ++j) csum[j] += g8[j];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) part[rg * K + c * 8 + j] = csum[j];
            named_sync(1, 128);
            // prefix of the earlier row groups + chunk total
            float G[8], tot[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { G[j] = 0.f; tot[j] = 0.f; }
#pragma unroll
            for (int r2 = 0; r2 < NRG; ++r2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float x = part[r2 * K + c * 8 + j];
                    tot[j] += x;
                    if (r2 < rg) G[j] += x;
                }
            }
            named_sync(1, 128);                    // `part` may be rewritten by the next item from here on
            if (rg == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) dvec[(n & 3) * K + c * 8 + j] = __expf(tot[j]);
            }
            // from here on G is kept in log2 units so every exponential is a single ex2
#pragma unroll
            for (int j = 0; j < 8; ++j) G[j] *= 1.44269504088896340736f;
            // pass B: walk the rows; rows past T were zero-filled by the loaders (gk = 0, q = k = 0)
#pragma unroll 4
            for (int i = 0; i < RPG; ++i) {
                const int r = rg * RPG + i;
                uint4 *qp4 = reinterpret_cast<uint4 *>(qk_tile + c * GQK + r * 16);
                uint4 *kp4 = reinterpret_cast<uint4 *>(qk_tile + c * GQK + (64 + r) * 16);
                float g8[8], q8[8], k8[8];
                unpack8(*reinterpret_cast<const uint4 *>(g_tile + c * GG + r * 16), g8);
                unpack8(*qp4, q8);
                unpack8(*kp4, k8);
                uint32_t qp[4], kp[4];
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    G[j] = fmaf(g8[j], 1.44269504088896340736f, G[j]);
                    G[j + 1] = fmaf(g8[j + 1], 1.44269504088896340736f, G[j + 1]);
                    const float e0 = ex2_approx(G[j]) * scale, e1 = ex2_approx(G[j + 1]) * scale;
                    const float i0 = ex2_approx(-G[j]), i1 = ex2_approx(-G[j + 1]);
                    qp[j >> 1] = pack_bf16(q8[j] * e0, q8[j + 1] * e1);
                    kp[j >> 1] = pack_bf16(k8[j] * i0, k8[j + 1] * i1);
                }
                *qp4 = make_uint4(qp[0], qp[1], qp[2], qp[3]);
                *kp4 = make_uint4(kp[0], kp[1], kp[2], kp[3]);
            }
            fence_proxy_async_smem();
            mbar_arrive(&bars[B_QK_FULL0 + s]);
            mbar_arrive(&bars[B_G_EMPTY0 + s]);
            if (p == 0) TRACE(0, n, 1);
        }
    } else if (warp == 5 || warp == 6) {
        // ====================== loaders: raw q, k -> operand tile of the stage, gk -> side tile ======================
        const int s = warp - 5;                    // warp 5: even items (stage 0), warp 6: odd items (stage 1)
        constexpr uint32_t GG = cfg::GG;
        const uint32_t qk_tile = smem_u32(smem + cfg::OFF_QK + s * cfg::QK_BYTES);
        const uint32_t g_tile = smem_u32(smem + cfg::OFF_G + s * cfg::G_BYTES);
        for (int n = s; n < n_items; n += 2) {
            const int t0 = n * C;
            wait_bar(&bars[B_QK_EMPTY0 + s], ((n >> 1) & 1) ^ 1);
            wait_bar(&bars[B_G_EMPTY0 + s], ((n >> 1) & 1) ^ 1);
            if (lane == 0) TRACE(1, n, 0);
            for (int cc = lane; cc < KC; cc += 32) {
#pragma unroll 8
                for (int r = 0; r < C; ++r) {
                    const int t = t0 + r;
                    const uint32_t nb = t < T ? 16u : 0u;
                    const size_t off = qbase + (size_t)(t < T ? t : 0) * K + cc * 8;
                    cp_async16(qk_tile + cc * GQK + r * 16, q + off, nb);
                    cp_async16(qk_tile + cc * GQK + (64 + r) * 16, k + off, nb);
                    cp_async16(g_tile + cc * GG + r * 16, gk + off, nb);
                }
            }
            if (lane == 0) TRACE(1, n, 1);
            cp_async_wait_all();
            mbar_arrive(&bars[B_RAW_FULL0 + s]);
            if (lane == 0) TRACE(1, n, 2);
        }
    } else if (warp == 7) {
        // ====================== loader: v tile (single stage) ======================
        const uint32_t v_tile = smem_u32(smem + cfg::OFF_V);
        for (int n = 0; n < n_items; ++n) {
            const int t0 = n * C;
            wait_bar(&bars[B_V_EMPTY], (n & 1) ^ 1);
#pragma unroll 8
            for (int i = 0; i < 32; ++i) {
                const int j = lane + 32 * i;
                const int sr = j >> 4, vc = j & 15;
                const int t = t0 + sr;
                cp_async16(v_tile + vc * GV + sr * 16, v + vbase + (size_t)(t < T ? t : 0) * V + v0 + vc * 8, t < T ? 16u : 0u);
            }
            cp_async_wait_all();
            fence_proxy_async_smem();
            mbar_arrive(&bars[B_V_FULL]);
        }
    } else if (warp == 4) {
        // ====================== MMA issuer (one thread) ======================
        if (lane == 0) {
            const uint32_t id_p = idesc_bf16(128, 64, 0, 0);     // (0),(1): K-major x K-major, N = 64
            const uint32_t id_o = idesc_bf16(128, 64, 1, 0);     // (2): MN-major A (v^T), K-major B (P)
            const uint32_t id_s = idesc_bf16(128, K, 1, 1);      // (3): MN-major A (v^T), MN-major B (k~)
            const uint32_t p_tile = smem_u32(smem + cfg::OFF_P);
            for (int n = 0; n < n_items; ++n) {
                const int s = n & 1;
                const uint32_t qk_tile = smem_u32(smem + cfg::OFF_QK + s * cfg::QK_BYTES);
                const uint32_t v_tile = smem_u32(smem + cfg::OFF_V);
                wait_bar(&bars[B_QK_FULL0 + s], (n >> 1) & 1);
                wait_bar(&bars[B_P_TEMPTY], (n & 1) ^ 1);
                tc_fence_after();
                TRACE(2, n, 0);
                // (0) P = [q~;k~] k~^T
#pragma unroll 4
                for (int ks = 0; ks < K / 16; ++ks) {
                    const uint64_t ad = smem_desc(qk_tile + ks * 2 * GQK, GQK, 128);
                    const uint64_t bd = smem_desc(qk_tile + ks * 2 * GQK + 64 * 16, GQK, 128);
                    mma_ss(tmem + COL_P, ad, bd, id_p, ks > 0);
                }
                mma_commit(&bars[B_P_FULL]);
                wait_bar(&bars[B_SA_FULL], n & 1);
                wait_bar(&bars[B_O_EMPTY], (n & 1) ^ 1);
                tc_fence_after();
                TRACE(2, n, 1);
                // (1) OT = SA q~^T   (A from TMEM)
#pragma unroll 4
                for (int ks = 0; ks < K / 16; ++ks) {
                    const uint64_t bd = smem_desc(qk_tile + ks * 2 * GQK, GQK, 128);
                    mma_ts(tmem + COL_OT, tmem + COL_SA + ks * 8, bd, id_p, ks > 0);
                }
                wait_bar(&bars[B_PS_FULL], n & 1);
                wait_bar(&bars[B_V_FULL], n & 1);
                tc_fence_after();
                TRACE(2, n, 2);
                // (2) OT += v^T P^T
#pragma unroll
                for (int ks = 0; ks < C / 16; ++ks) {
                    const uint64_t ad = smem_desc(v_tile + ks * 256, 128, GV);
                    const uint64_t bd = smem_desc(p_tile + ks * 2 * GP, GP, 128);
                    mma_ss(tmem + COL_OT, ad, bd, id_o, 1);
                }
                mma_commit(&bars[B_O_FULL]);
                mma_commit(&bars[B_PS_EMPTY]);
                // (3) ST += v^T k~
#pragma unroll
                for (int ks = 0; ks < C / 16; ++ks) {
                    const uint64_t ad = smem_desc(v_tile + ks * 256, 128, GV);
                    const uint64_t bd = smem_desc(qk_tile + 64 * 16 + ks * 256, 128, GQK);
                    mma_ss(tmem + COL_ST, ad, bd, id_s, 1);
                }
                mma_commit(&bars[B_ST_FULL]);
                mma_commit(&bars[B_QK_EMPTY0 + s]);
                mma_commit(&bars[B_V_EMPTY]);
                TRACE(2, n, 3);
            }
        }
        __syncwarp();
    } else if (warp >= 16) {
        // ====================== warps 16,17: causal mask of P (rows t = TMEM lanes 0..63) ======================
        const int qd = warp - 16, r = qd * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
        uint8_t *p_tile = smem + cfg::OFF_P;
        for (int n = 0; n < n_items; ++n) {
            wait_bar(&bars[B_P_FULL], n & 1);
            tc_fence_after();
            if (r == 0) TRACE(3, n, 0);
            uint32_t pr[2][32];
            tmem_ld32(tmem + lane_addr + COL_P, pr[0]);
            tmem_ld32(tmem + lane_addr + COL_P + 32, pr[1]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars[B_P_TEMPTY]);
            wait_bar(&bars[B_PS_EMPTY], (n & 1) ^ 1);
            // row t = r ; keep s <= t ; bf16 ; P tile [s/8][t][8]
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int s0 = g * 8 + 2 * j;
                    const float a = s0 <= r ? __uint_as_float(pr[s0 >> 5][s0 & 31]) : 0.f;
                    const float b = s0 + 1 <= r ? __uint_as_float(pr[(s0 + 1) >> 5][(s0 + 1) & 31]) : 0.f;
                    w[j] = pack_bf16(a, b);
                }
                *reinterpret_cast<uint4 *>(p_tile + g * GP + r * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            fence_proxy_async_smem();
            mbar_arrive(&bars[B_PS_FULL]);
            if (r == 0) TRACE(3, n, 1);
        }
    } else if (warp >= 8 && warp < 12) {
        // ====================== warps 8-11: output epilogue  OT[lane = v][col = t] -> o[t][v0 + lane] ======================
        const int qd = warp - 8, r = qd * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
        for (int n = 0; n < n_items; ++n) {
            const int t0 = n * C;
            wait_bar(&bars[B_O_FULL], n & 1);
            tc_fence_after();
            if (r == 0) TRACE(4, n, 0);
            uint32_t orr[2][32];
            tmem_ld32(tmem + lane_addr + COL_OT, orr[0]);
            tmem_ld32(tmem + lane_addr + COL_OT + 32, orr[1]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars[B_O_EMPTY]);
            bf16 *ob = o + vbase + (size_t)t0 * V + v0 + r;
            const int nrow = min(C, T - t0);
            if (nrow == C) {
#pragma unroll
                for (int t = 0; t < C; ++t) { *ob = __float2bfloat16_rn(__uint_as_float(orr[t >> 5][t & 31])); ob += V; }
            } else {
#pragma unroll
                for (int t = 0; t < C; ++t) {
                    if (t < nrow) ob[(size_t)t * V] = __float2bfloat16_rn(__uint_as_float(orr[t >> 5][t & 31]));
                }
            }
            if (r == 0) TRACE(4, n, 1);
        }
    } else if (warp >= 12 && warp < 16) {
        // ====================== WG3: state pass ======================
        const int qd = warp - 12, r = qd * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
        const size_t sbase = (size_t)bh * K * V + v0 + r;          // + kappa * V
        // initial state -> ST (fp32) and SA (bf16)
#pragma unroll 1
        for (int cb = 0; cb < K / 32; ++cb) {
            uint32_t f[32], pk[16];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float x = h0 != nullptr ? load_dyn(h0, h0_dtype, sbase + (size_t)(cb * 32 + j) * V) : 0.f;
                f[j] = __float_as_uint(x);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(__uint_as_float(f[2 * j]), __uint_as_float(f[2 * j + 1]));
            tmem_st32(tmem + lane_addr + COL_ST + cb * 32, f);
            tmem_st16(tmem + lane_addr + COL_SA + cb * 16, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars[B_SA_FULL]);
        for (int n = 0; n < n_items; ++n) {
            const int s = n & 1;
            wait_bar(&bars[B_QK_FULL0 + s], (n >> 1) & 1);         // dvec of this item is published with it
            wait_bar(&bars[B_ST_FULL], n & 1);
            tc_fence_after();
            if (r == 0) TRACE(5, n, 0);
            const float *dv = dvec + (n & 3) * K;
            const bool last = n == n_items - 1;
#pragma unroll 1
            for (int cb = 0; cb < K / 32; ++cb) {
                uint32_t f[32], pk[16];
                tmem_ld32(tmem + lane_addr + COL_ST + cb * 32, f);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 d4 = *reinterpret_cast<const float4 *>(dv + cb * 32 + j);
                    f[j + 0] = __float_as_uint(__uint_as_float(f[j + 0]) * d4.x);
                    f[j + 1] = __float_as_uint(__uint_as_float(f[j + 1]) * d4.y);
                    f[j + 2] = __float_as_uint(__uint_as_float(f[j + 2]) * d4.z);
                    f[j + 3] = __float_as_uint(__uint_as_float(f[j + 3]) * d4.w);
                }
                if (!last) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(__uint_as_float(f[2 * j]), __uint_as_float(f[2 * j + 1]));
                    tmem_st32(tmem + lane_addr + COL_ST + cb * 32, f);
                    tmem_st16(tmem + lane_addr + COL_SA + cb * 16, pk);
                } else if (ht != nullptr) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) ht[sbase + (size_t)(cb * 32 + j) * V] = __uint_as_float(f[j]);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars[B_SA_FULL]);
            if (r == 0) TRACE(5, n, 1);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int K>
int launch(const void *q, const void *k, const void *v, const void *gk, const void *h0, int h0_dtype, void *o,
           float *ht, int B, int H, int T, int V, float scale, cudaStream_t st, long long *trace = nullptr) {
    using cfg = Cfg<K>;
    static thread_local bool configured = false;
    if (!configured) {
        LINA_CUDA_OK(cudaFuncSetAttribute(gla_chunk_fwd_sm100_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)cfg::SMEM));
        configured = true;
    }
    dim3 grid(V / BV, B * H);
    gla_chunk_fwd_sm100_kernel<K><<<grid, NTHREADS, cfg::SMEM, st>>>(
        (const bf16 *)q, (const bf16 *)k, (const bf16 *)v, (const bf16 *)gk, h0, h0_dtype, (bf16 *)o, ht, T, V, scale, trace);
    LINA_LAUNCH_OK("gla_chunk_fwd_sm100_kernel");
    return LINA_OK;
}

bool tc_eligible(int B, int H, int T, int K, int V, int dtype) {
    return dtype == LINA_BF16 && (K == 64 || K == 128 || K == 256) && V % BV == 0 && T >= 32 &&
           (long long)B * H <= 65535;
}

}  // namespace

extern "C" int lina_gla_chunk_fwd_uses_tensor_cores(int B, int H, int T, int K, int V, int dtype) {
    return tc_eligible(B, H, T, K, V, dtype) ? 1 : 0;
}

extern "C" size_t lina_gla_chunk_fwd_workspace_bytes(int B, int H, int T, int K, int V, int dtype) {
    (void)B; (void)H; (void)T; (void)K; (void)V; (void)dtype;
    return 16;
}

extern "C" int lina_gla_chunk_fwd(const void *q, const void *k, const void *v, const void *gk, const void *h0,
                                  int h0_dtype, void *o, float *ht, void *ws, int B, int H, int T, int K, int V,
                                  int dtype, float scale, void *stream) {
    (void)ws;
    if (!tc_eligible(B, H, T, K, V, dtype))
        return lina_gla_recurrent_fwd_impl(q, k, v, gk, h0, h0_dtype, o, ht, B, H, T, K, V, dtype, scale, stream);
    LINA_REQUIRE(q && k && v && gk && o, LINA_ERR_BAD_ARG, "gla_chunk_fwd: null tensor pointer");
    LINA_REQUIRE(h0 == nullptr || lina_dtype_ok(h0_dtype), LINA_ERR_BAD_ARG, "gla_chunk_fwd: bad h0 dtype");
    cudaStream_t st = (cudaStream_t)stream;
    if (K == 64) return launch<64>(q, k, v, gk, h0, h0_dtype, o, ht, B, H, T, V, scale, st);
    if (K == 128) return launch<128>(q, k, v, gk, h0, h0_dtype, o, ht, B, H, T, V, scale, st);
    return launch<256>(q, k, v, gk, h0, h0_dtype, o, ht, B, H, T, V, scale, st);
}

// bring-up: same kernel with a clock64 timeline of CTA (0,0): trace[6 roles][64 items][4 events]
extern "C" int lina_debug_gla_chunk_trace(const void *q, const void *k, const void *v, const void *gk, void *o, int B,
                                          int H, int T, int K, int V, float scale, long long *trace, void *stream) {
    LINA_REQUIRE(tc_eligible(B, H, T, K, V, LINA_BF16) && K == 256, LINA_ERR_UNSUPPORTED, "trace: K=256 bf16 only");
    return launch<256>(q, k, v, gk, nullptr, 0, o, nullptr, B, H, T, V, scale, (cudaStream_t)stream, trace);
}
