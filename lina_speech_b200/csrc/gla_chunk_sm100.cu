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
// Shared-memory operands are 128-byte-swizzled [rows][64] blocks (TMA's CU_TENSOR_MAP_SWIZZLE_128B = UMMA
// SWIZZLE_128B, see sm100.cuh): the raw q, k, gk, v tiles are landed by TMA directly in operand layout, q and k
// are rescaled IN PLACE, and the same k~ bytes serve as K-major operand of (0) and MN-major operand of (3), the
// same v bytes as the MN-major A of (2) and (3).
//
// Warp roles (640 threads):
//   warp 19     TMA loader (one thread): 3-D tensor maps (K|V, T, B*H), 64x64 boxes, rows past T zero-filled;
//   warps 0-7   gate pre-pass IN PLACE on the landed stage: cumsum over the chunk, q -> q~, k -> k~ (two stages);
//   warp 18     MMA issuer (one thread);
//   warps 16,17 causal mask (P: TMEM -> bf16 smem);      warps 8-11 output epilogue (OT: TMEM -> global);
//   warps 12-15 state pass (ST *= exp(G_C), refresh SA, final state).
// Synchronisation is mbarrier-only; global-load latency is hidden by TMA running a stage ahead.
//
// HBM traffic per CTA = q,k,gk once + its v slice + its o slice; q,k,gk are re-read by the V/128 CTAs of
// the same (b,h) (L2 hits).  Algorithmic bytes per token per head: (3K + 2V) * 2.
#include "common.cuh"
#include "sm100.cuh"
#include "tma.cuh"

using namespace sm100;

extern int g_lina_variant[16];

int lina_gla_recurrent_fwd_impl(const void *q, const void *k, const void *v, const void *gk, const void *h0,
                                int h0_dtype, void *o, float *ht, int B, int H, int T, int K, int V, int dtype,
                                float scale, void *stream);

namespace {

constexpr int C = 64;            // chunk length (tokens)
constexpr int BV = 128;          // V slice per CTA
constexpr int NTHREADS = 640;    // 20 warps, see the role table above
constexpr int NPREP = 256;       // warps 0-7
constexpr uint32_t VT_BYTES = 2 * 8192;   // v tile : 2 blocks of [64 rows s][64 v] (128 B rows, 128B-swizzled)
constexpr uint32_t PT_BYTES = 8192;       // P tile : [64 rows t][64 s]
// TMEM columns
constexpr uint32_t COL_ST = 0, COL_OT = 256, COL_P = 320, COL_SA = 384;

template <int K> struct Cfg {
    static constexpr int KC = K / 8;                       // 16-byte groups along K
    static constexpr int KB = K / 64;                      // 64-wide (128-byte) swizzle blocks along K
    static constexpr uint32_t QK_BLK = 128 * 128;          // [128 rows: q~ 0..63, k~ 64..127][64 k]
    static constexpr uint32_t QK_BYTES = KB * QK_BLK;
    static constexpr uint32_t G_BLK = 64 * 128;            // [64 rows][64 k]
    static constexpr uint32_t G_BYTES = KB * G_BLK;
    static constexpr int NRG = NPREP / KC;                 // row groups of the pre-pass
    static constexpr int RPG = C / NRG;                    // rows per group
    static constexpr uint32_t OFF_QK = 0;
    static constexpr uint32_t OFF_G = OFF_QK + 2 * QK_BYTES;
    static constexpr uint32_t OFF_V = OFF_G + 2 * G_BYTES;
    static constexpr uint32_t OFF_P = OFF_V + VT_BYTES;               // single v stage
    static constexpr uint32_t OFF_DVEC = OFF_P + PT_BYTES;            // [3][K] fp32
    static constexpr uint32_t OFF_PART = OFF_DVEC + 3 * K * 4;        // [NRG-1][K] fp32
    static constexpr uint32_t OFF_BAR = OFF_PART + (NRG - 1) * K * 4; // mbarriers
    static constexpr uint32_t SMEM = OFF_BAR + 32 * 8 + 16;
    static constexpr uint32_t RAW_TX = 3u * K * C * 2u;               // bytes landed per stage by TMA (q, k, gk)
    static_assert(SMEM <= 232448, "shared memory budget");
    static_assert(OFF_G % 1024 == 0 && OFF_V % 1024 == 0 && OFF_P % 1024 == 0, "swizzled tiles need 1024-byte alignment");
};

enum { B_QK_FULL0 = 0, B_QK_FULL1, B_QK_EMPTY0, B_QK_EMPTY1, B_P_FULL, B_P_TEMPTY, B_PS_FULL, B_PS_FULL1, B_PS_EMPTY, B_PS_EMPTY1,
       B_O_FULL, B_O_EMPTY, B_ST_FULL, B_RAW_FULL0, B_RAW_FULL1, B_G_EMPTY0, B_G_EMPTY1, B_V_FULL0, B_V_FULL1,
       B_V_EMPTY0, B_V_EMPTY1, B_SA_BLK0 /* .. B_SA_BLK0 + K/32 - 1: one per 32-column block of the state */, B_COUNT = B_SA_BLK0 + 8 };
static_assert(B_COUNT <= 32, "mbarrier slots");

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
#ifdef LINA_GLA_TRACE        // liblina_b200_debug.so only: the product kernel carries no timeline code
#define TRACE(role, n, ev) do { if (trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && (n) < TR_MAXN) \
    trace[((role) * TR_MAXN + (n)) * TR_EV + (ev)] = clock64(); } while (0)
#else
#define TRACE(role, n, ev) do { } while (0)
#endif

__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void unpack8(const uint4 &raw, float *f) {
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

struct TMaps { CUtensorMap q, k, g, v; };

// Work list of the PAIR variant (OPT bit 8).  The grid is 1-D, two CTAs (one cluster) per entry.  `U` tiles = (batch, head) x
// pairs of V slices; entry ci < rem: FIRST part (items [0, n_split)) of tile ci; rem <= ci < U: all of tile ci; ci >= U: SECOND
// part of tile ci - U.  A first part leaves its state slice in hx[(tile * 2 + rank)][K][128] (fp32) and raises flags[tile * 2 + rank];
// the second part starts from it.  Blocks are dispatched in index order, so a first part has long finished when the second
// part of the same tile (U entries later) starts: the cut tiles fill what would be a quarter-empty last wave.
struct SplitArgs { float *hx; uint32_t *flags; int rem, n_split, U, nvp; };

// OPT bit 1: the gate pre-pass keeps its gk rows in registers between the column-sum pass and the rescale pass
//            (8 fewer 16-byte shared loads per thread and item).
// OPT bit 2 (PRE): the operands arrive PRE-GATED -- tm.q / tm.k map q~ = scale q e^G and k~ = k e^-G (written by
//            lina_gla_prefill_prep_gated together with `decay` [B,H,NT,K] = e^{G_C}); TMA lands them straight in operand
//            layout, the gate pre-pass warps idle, gk is never read, the chunk decay vector is a 1-D bulk copy.
template <int K, int OPT>
__global__ void __launch_bounds__(NTHREADS, 1)
gla_chunk_fwd_sm100_kernel(const __grid_constant__ TMaps tm, const void *__restrict__ h0, int h0_dtype,
                           bf16 *__restrict__ o, float *__restrict__ ht, int T, int V, int H, int bthd, float scale,
                           long long *__restrict__ trace, const float *__restrict__ decay, const SplitArgs sp) {
    constexpr bool PRE = (OPT & 4) != 0;
    // OPT bit 3 (ROW, with PRE): the chunk decay scales the VALUE dim (rows of the transposed state = TMEM lanes) instead of
    //            the key dim: decay is [B,H,NT,V].  This is the form the backward needs (contraction over V, state S^T).
    // OPT bit 4 (OUT32): o is written as fp32 (the backward's dq~ / dk~ partial sums feed a cumsum).
    constexpr bool ROW = (OPT & 8) != 0, OUT32 = (OPT & 16) != 0;
    // OPT bit 5 (STATE2, with PRE): the idle pre-pass warps 0-3 and 4-7 join warps 12-15 in the state pass (a warp may touch
    //            TMEM lanes 32*(warp%4)..+31 only, so they mirror warps 12-15 lane for lane).  The state is handled in
    //            32-column blocks, block b by group b % NG, each block with its own mbarrier so that the q~ S MMA of the
    //            next item starts on block 0 while the later blocks are still being rescaled.
    constexpr bool STATE2 = (OPT & 32) != 0;
    static_assert(!STATE2 || PRE, "the extra state-pass warpgroups are the pre-pass warps");
    // (OPT bits 6 / 7 were a cluster-multicast operand load and a three-stage operand ring: measured neutral / slower in round 1
    //  -- DESIGN 4.1b -- and removed.)
    // OPT bit 8 (PAIR, with PRE | STATE2): two CTAs holding neighbouring V slices of one (batch, head) form a CLUSTER and share the
    //            score MMA: P = q~ k~^T does not depend on the V slice, so CTA r computes it for the items n with n % 2 == r, masks
    //            it, and delivers the bf16 tile to both CTAs (own shared memory + one 8 KB DSMEM bulk copy that completes on the
    //            peer's mbarrier).  Per item the tensor pipe of a CTA then runs 16 + 4 + 4 (+ 16 every other item) MMAs instead
    //            of 40.  Issue order per item: (1) OT = SA q~^T, (3) ST += v^T k~, (2) OT += v^T P^T, (0) of the NEXT item when it
    //            is this CTA's -- the state pass (the serial chain of the kernel) starts as early as possible and (2), (0) run
    //            under it.  The grid is a work list (SplitArgs) so that tiles can be cut in two along T.
    constexpr bool PAIR = (OPT & 256) != 0;
    static_assert(!PAIR || (PRE && (OPT & 32) != 0 && (OPT & (8 | 16)) == 0), "PAIR builds on the pre-gated, three-warpgroup variant");
    constexpr int NB = K / 32;
    constexpr int NG = STATE2 ? (NB >= 3 ? 3 : NB) : 1;
    static_assert(!ROW || PRE, "row decay needs pre-gated operands");
    using cfg = Cfg<K>;
    // With pre-gated operands the gk side tiles (2 * G_BYTES = QK_BYTES) are unused: they hold a second v stage (16 KB) and, with
    // PAIR, the second P tile (8 KB).
    constexpr int NS = 2;
    constexpr bool V2 = PRE;
    constexpr uint32_t OFF_P1 = cfg::OFF_G + VT_BYTES;            // PAIR: P tile of the odd items
    static_assert(!V2 || 2 * cfg::G_BYTES >= VT_BYTES, "the second v stage lives in the gk tiles");
    static_assert(!PAIR || 2 * cfg::G_BYTES >= VT_BYTES + PT_BYTES, "second v stage + second P tile live in the gk tiles");
    constexpr uint32_t V_STAGE1 = V2 ? cfg::OFF_G : cfg::OFF_V;
    constexpr int KC = cfg::KC, KB = cfg::KB, NRG = cfg::NRG, RPG = cfg::RPG;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + cfg::OFF_BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + cfg::OFF_BAR + 32 * 8);
    float *dvec = reinterpret_cast<float *>(smem + cfg::OFF_DVEC);
    float *part = reinterpret_cast<float *>(smem + cfg::OFF_PART);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
    const int n_total = (T + C - 1) / C;
    int bh = blockIdx.y, v0 = blockIdx.x * BV, n_begin = 0, n_items = n_total, piece = 0, hx_slot = 0;
    if (PAIR) {
        const int ci = blockIdx.x >> 1;
        int tile = ci;
        if (ci < sp.rem) piece = 1;
        else if (ci >= sp.U) { tile = ci - sp.U; piece = 2; }
        bh = tile / sp.nvp;
        v0 = ((tile - bh * sp.nvp) * 2 + (int)crank) * BV;
        if (piece == 1) n_items = sp.n_split;
        else if (piece == 2) { n_begin = sp.n_split; n_items = n_total - sp.n_split; }
        hx_slot = tile * 2 + (int)crank;
    }
    const int bb = bh / H, hh = bh - bb * H;
    // o is [B,H,T,V] (bthd == 0) or [B,T,H,V] (bthd == 1)
    const size_t obase = bthd ? ((size_t)bb * T * H + hh) * V : (size_t)bh * T * V;
    const size_t o_tstride = bthd ? (size_t)H * V : (size_t)V;

    if (tid == 0) {
        if (smem_u32(smem) & 1023u) { printf("gla_chunk_sm100: dynamic smem base not 1024-byte aligned\n"); __trap(); }
        mbar_init(&bars[B_QK_FULL0], NPREP); mbar_init(&bars[B_QK_FULL1], NPREP);
        mbar_init(&bars[B_QK_EMPTY0], 1); mbar_init(&bars[B_QK_EMPTY1], 1);
        mbar_init(&bars[B_P_FULL], 1); mbar_init(&bars[B_P_TEMPTY], 64);
        if (PAIR) {
            // P tile j is written by CTA j: 64 mask threads arrive at home, one expect_tx arrive + 8 KB of bulk copy at the peer;
            // it is free again when the (2) MMAs of BOTH CTAs have read it (multicast tcgen05.commit: 2 arrivals)
            mbar_init(&bars[B_PS_FULL], crank == 0 ? 64 : 1); mbar_init(&bars[B_PS_FULL1], crank == 1 ? 64 : 1);
            mbar_init(&bars[B_PS_EMPTY], 2); mbar_init(&bars[B_PS_EMPTY1], 2);
        } else {
            mbar_init(&bars[B_PS_FULL], 64); mbar_init(&bars[B_PS_EMPTY], 1);
        }
        mbar_init(&bars[B_O_FULL], 1); mbar_init(&bars[B_O_EMPTY], 128);
        mbar_init(&bars[B_ST_FULL], 1);
        for (int b = 0; b < NB; ++b) mbar_init(&bars[B_SA_BLK0 + b], 128);
        mbar_init(&bars[B_RAW_FULL0], 1); mbar_init(&bars[B_RAW_FULL1], 1);
        mbar_init(&bars[B_G_EMPTY0], NPREP); mbar_init(&bars[B_G_EMPTY1], NPREP);
        mbar_init(&bars[B_V_FULL0], 1); mbar_init(&bars[B_V_EMPTY0], 1);
        mbar_init(&bars[B_V_FULL1], 1); mbar_init(&bars[B_V_EMPTY1], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();        // both CTAs' mbarriers are initialised before the peer signals them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    auto qk_stage_off = [](int st) -> uint32_t { return cfg::OFF_QK + st * cfg::QK_BYTES; };
    // chunk-decay ring: NS + 1 slots (the loader runs NS items ahead of the state pass); the 4th lives in the unused `part`
    auto dvec_slot = [&](int n) -> float * { const int sl = n % (NS + 1); return sl < 3 ? dvec + sl * K : part; };
    int sg = -1;                                   // state-pass group of this warp
    if (warp >= 12 && warp < 16) sg = 0;
    else if (NG > 1 && warp < 4) sg = 1;
    else if (NG > 2 && warp >= 4 && warp < 8) sg = 2;
    if (warp < 8 && sg < 0) {
      if (!PRE) {
        // ====================== warps 0-7: gate pre-pass, in place on the landed q / k rows ======================
        const int p = tid;                         // 0..255
        const int c = p % KC, rg = p / KC;         // 16-byte column group, row group
        const uint32_t cb = (uint32_t)(c >> 3), c16 = (uint32_t)(c & 7);
        for (int n = 0; n < n_items; ++n) {
            const int s = n & 1;
            uint8_t *qk_tile = smem + cfg::OFF_QK + s * cfg::QK_BYTES + cb * cfg::QK_BLK;
            const uint8_t *g_tile = smem + cfg::OFF_G + s * cfg::G_BYTES + cb * cfg::G_BLK;
            wait_bar(&bars[B_RAW_FULL0 + s], (n >> 1) & 1);
            if (p == 0) TRACE(0, n, 0);
            // pass A: column sums of gk over this thread's rows
            float csum[8];
            uint4 graw[(OPT & 2) ? RPG : 1];
#pragma unroll
            for (int j = 0; j < 8; ++j) csum[j] = 0.f;
#pragma unroll
            for (int i = 0; i < RPG; ++i) {
                float g8[8];
                const uint4 gr = *reinterpret_cast<const uint4 *>(g_tile + sw128_off(rg * RPG + i, c16));
                if (OPT & 2) graw[i] = gr;
                unpack8(gr, g8);
#pragma unroll
                for (int j = 0; j < 8; ++j) csum[j] += g8[j];
            }
            if (rg < NRG - 1) {
#pragma unroll
                for (int j = 0; j < 8; j += 4)
                    *reinterpret_cast<float4 *>(part + rg * K + c * 8 + j) = make_float4(csum[j], csum[j + 1], csum[j + 2], csum[j + 3]);
            }
            named_sync(1, NPREP);
            // prefix of the earlier row groups (+ chunk total in the last row group)
            float G[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) G[j] = 0.f;
            for (int r2 = 0; r2 < rg; ++r2) {
#pragma unroll
                for (int j = 0; j < 8; j += 4) {
                    const float4 x = *reinterpret_cast<const float4 *>(part + r2 * K + c * 8 + j);
                    G[j] += x.x; G[j + 1] += x.y; G[j + 2] += x.z; G[j + 3] += x.w;
                }
            }
            named_sync(1, NPREP);                  // `part` may be rewritten by the next item from here on
            if (rg == NRG - 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) dvec_slot(n)[c * 8 + j] = __expf(G[j] + csum[j]);
            }
            // from here on G is kept in log2 units so every exponential is a single ex2
#pragma unroll
            for (int j = 0; j < 8; ++j) G[j] *= 1.44269504088896340736f;
            // pass B: walk the rows; rows past T were zero-filled by TMA (gk = 0, q = k = 0)
#pragma unroll
            for (int i = 0; i < RPG; ++i) {
                const int r = rg * RPG + i;
                uint4 *qp4 = reinterpret_cast<uint4 *>(qk_tile + sw128_off(r, c16));
                uint4 *kp4 = reinterpret_cast<uint4 *>(qk_tile + sw128_off(64 + r, c16));
                float g8[8], q8[8], k8[8];
                if (OPT & 2) unpack8(graw[i], g8);
                else unpack8(*reinterpret_cast<const uint4 *>(g_tile + sw128_off(r, c16)), g8);
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
      }
    } else if (warp == 19) {
        // ====================== TMA loader (one elected thread): q, k -> operand tile of the stage, gk -> side tile, v ======================
        if (elect_one_sync()) {
            tma_prefetch_desc(&tm.q); tma_prefetch_desc(&tm.k); tma_prefetch_desc(&tm.v);
            if (!PRE) tma_prefetch_desc(&tm.g);
            // Two independent streams (q~/k~[/gk] stages and v tiles) served by ONE thread: it polls both EMPTY barriers
            // instead of blocking on one, so a full v ring never stops the operand ring from running ahead (and vice versa).
            int nq = 0, nv = 0;
            long long t_idle = clock64();
            while (nq < n_items || nv < n_items) {
                bool progressed = false;
                if (nq < n_items) {
                    const int n = nq, s = n % NS, u = n / NS, t0 = (n_begin + n) * C;
                    if (mbar_try_wait(&bars[B_QK_EMPTY0 + s], (u & 1) ^ 1) &&
                        (PRE || mbar_try_wait(&bars[B_G_EMPTY0 + s], (u & 1) ^ 1))) {
                        const uint32_t qk_tile = smem_u32(smem + qk_stage_off(s));
                        const uint32_t g_tile = smem_u32(smem + cfg::OFF_G + s * cfg::G_BYTES);
                        TRACE(1, n, 0);
                        mbar_expect_tx(&bars[B_RAW_FULL0 + s], PRE ? (2u * K * C * 2u + (ROW ? 0u : K * 4u)) : cfg::RAW_TX);
                        if (PRE && !ROW)   // chunk-decay ring slot <- decay[b, h, n, :]
                            tma_load_1d(smem_u32(dvec_slot(n)), decay + ((size_t)bh * n_total + n_begin + n) * K, K * 4u,
                                        &bars[B_RAW_FULL0 + s]);
#pragma unroll
                        for (int kb = 0; kb < KB; ++kb) {
                            if (!PRE) tma_load_4d(g_tile + kb * cfg::G_BLK, &tm.g, kb * 64, t0, hh, bb, &bars[B_RAW_FULL0 + s]);
                            tma_load_4d(qk_tile + kb * cfg::QK_BLK, &tm.q, kb * 64, t0, hh, bb, &bars[B_RAW_FULL0 + s]);
                            tma_load_4d(qk_tile + kb * cfg::QK_BLK + 8192, &tm.k, kb * 64, t0, hh, bb, &bars[B_RAW_FULL0 + s]);
                        }
                        TRACE(1, n, 1);
                        ++nq;
                        progressed = true;
                    }
                }
                if (nv < n_items) {
                    const int n = nv, t0 = (n_begin + n) * C;
                    const int sv = V2 ? (n & 1) : 0, uv = V2 ? (n >> 1) : n;         // v stage and its use count
                    if (mbar_try_wait(&bars[B_V_EMPTY0 + sv], (uv & 1) ^ 1)) {
                        mbar_expect_tx(&bars[B_V_FULL0 + sv], VT_BYTES);
                        const uint32_t v_tile = smem_u32(smem + (sv ? V_STAGE1 : cfg::OFF_V));
                        tma_load_4d(v_tile, &tm.v, v0, t0, hh, bb, &bars[B_V_FULL0 + sv]);
                        tma_load_4d(v_tile + 8192, &tm.v, v0 + 64, t0, hh, bb, &bars[B_V_FULL0 + sv]);
                        TRACE(1, n, 2);
                        ++nv;
                        progressed = true;
                    }
                }
                if (progressed) t_idle = clock64();
                else if (clock64() - t_idle > 4000000000LL) {
                    printf("gla_chunk_sm100: loader timeout (block %d,%d nq %d nv %d)\n", blockIdx.x, blockIdx.y, nq, nv);
                    __trap();
                }
            }
        }
        __syncwarp();
    } else if (warp == 18) {
        // ====================== MMA issuer (one elected thread) ======================
        if (elect_one_sync()) {
            const uint32_t id_p = idesc_bf16(128, 64, 0, 0);     // (0),(1): K-major x K-major, N = 64
            const uint32_t id_o = idesc_bf16(128, 64, 1, 0);     // (2): MN-major A (v^T), K-major B (P)
            const uint32_t id_s = idesc_bf16(128, K, 1, 1);      // (3): MN-major A (v^T), MN-major B (k~)
            const uint32_t p_tile = smem_u32(smem + cfg::OFF_P);
            const uint32_t v_tile = smem_u32(smem + cfg::OFF_V);
            // Descriptors are built ONCE; inside the loops only a compile-time constant is added to their 14-bit
            // (address >> 4) field (all operand addresses are < 256 KB, so the field never carries).  Rebuilding them per
            // instruction cost ~10 dependent uniform-datapath ops per MMA, and the single issuing thread is the kernel's
            // critical resource (40 MMAs per 64-token item).
            const uint32_t qk0 = smem_u32(smem + cfg::OFF_QK);
            const uint32_t qk1 = qk0 + cfg::QK_BYTES;
            const uint64_t d_q0 = smem_desc_sw128(qk0, 0, 1024), d_q1 = smem_desc_sw128(qk1, 0, 1024);
            const uint64_t d_k0 = smem_desc_sw128(qk0 + 8192, 0, 1024), d_k1 = smem_desc_sw128(qk1 + 8192, 0, 1024);
            const uint64_t d_k30 = smem_desc_sw128(qk0 + 8192, cfg::QK_BLK, 1024), d_k31 = smem_desc_sw128(qk1 + 8192, cfg::QK_BLK, 1024);
            const uint64_t d_v0 = smem_desc_sw128(v_tile, 8192, 1024), d_v1 = smem_desc_sw128(smem_u32(smem + V_STAGE1), 8192, 1024);
            const uint64_t d_p = smem_desc_sw128(p_tile, 0, 1024), d_p1 = smem_desc_sw128(smem_u32(smem + OFF_P1), 0, 1024);
            if (PAIR) {
                // (0) of item m (one of this CTA's): P = [q~;k~] k~^T into the CTA's own tensor-memory tile
                auto issue_scores = [&](int m) {
                    const int s = m & 1;
                    const uint64_t dq = s ? d_q1 : d_q0, dk = s ? d_k1 : d_k0;
                    wait_bar(&bars[B_RAW_FULL0 + s], (m >> 1) & 1);
                    wait_bar(&bars[B_P_TEMPTY], ((m >> 1) & 1) ^ 1);
                    tc_fence_after();
                    TRACE(2, m, 0);
#pragma unroll
                    for (int ks = 0; ks < K / 16; ++ks) {
                        const uint64_t off = (uint64_t)(((ks >> 2) * cfg::QK_BLK + (ks & 3) * 32) >> 4);
                        mma_ss(tmem + COL_P, dq + off, dk + off, id_p, ks > 0);
                    }
                    mma_commit(&bars[B_P_FULL]);
                };
                if (crank == 0) issue_scores(0);
                for (int n = 0; n < n_items; ++n) {
                    const int s = n & 1, u = n >> 1, pb = n & 1;
                    const uint64_t dq = s ? d_q1 : d_q0, dk3 = s ? d_k31 : d_k30, d_v = s ? d_v1 : d_v0;
                    if ((uint32_t)pb != crank) mbar_expect_tx(&bars[B_PS_FULL + pb], PT_BYTES);   // the peer's bulk copy of P_n lands here
                    wait_bar(&bars[B_RAW_FULL0 + s], u & 1);
                    wait_bar(&bars[B_O_EMPTY], (n & 1) ^ 1);
                    TRACE(0, n, 0);
                    // (1) OT = SA q~^T   (A from TMEM), two k-steps per 32-column block of the state as the blocks become ready
#pragma unroll
                    for (int cb = 0; cb < NB; ++cb) {
                        wait_bar(&bars[B_SA_BLK0 + cb], n & 1);
                        tc_fence_after();
                        if (cb == 0) TRACE(2, n, 1);
#pragma unroll
                        for (int ks = 2 * cb; ks < 2 * cb + 2; ++ks) {
                            const uint64_t off = (uint64_t)(((ks >> 2) * cfg::QK_BLK + (ks & 3) * 32) >> 4);
                            mma_ts(tmem + COL_OT, tmem + COL_SA + ks * 8, dq + off, id_p, ks > 0);
                        }
                    }
                    TRACE(0, n, 1);
                    wait_bar(&bars[B_V_FULL0 + s], u & 1);
                    tc_fence_after();
                    TRACE(0, n, 2);
                    // (3) ST += v^T k~
#pragma unroll
                    for (int ks = 0; ks < C / 16; ++ks)
                        mma_ss(tmem + COL_ST, d_v + (uint64_t)(ks * 128), dk3 + (uint64_t)(ks * 128), id_s, 1);
                    mma_commit(&bars[B_ST_FULL]);
                    mma_commit(&bars[B_QK_EMPTY0 + s]);           // q~_n, k~_n are done with: (0)_n, (1)_n, (3)_n
                    const bool next_mine = n + 1 < n_items && (uint32_t)((n + 1) & 1) == crank;
                    TRACE(0, n, 3);
                    wait_bar(&bars[B_PS_FULL + pb], u & 1);
                    fence_proxy_async_smem();
                    tc_fence_after();
                    TRACE(2, n, 2);
                    // (2) OT += v^T P^T
                    const uint64_t dp = pb ? d_p1 : d_p;
#pragma unroll
                    for (int ks = 0; ks < C / 16; ++ks)
                        mma_ss(tmem + COL_OT, d_v + (uint64_t)(ks * 128), dp + (uint64_t)(ks * 2), id_o, 1);
                    mma_commit(&bars[B_O_FULL]);
                    mma_commit_mc(&bars[B_PS_EMPTY + pb], (uint16_t)3);
                    mma_commit(&bars[B_V_EMPTY0 + s]);
                    // (0) of the next item, if it is ours: runs under the state pass of this one
                    if (next_mine) issue_scores(n + 1);
                    TRACE(2, n, 3);
                }
            } else {
            for (int n = 0; n < n_items; ++n) {
                const int s = n % NS, u = n / NS;
                const uint64_t dq = s == 0 ? d_q0 : d_q1, dk = s == 0 ? d_k0 : d_k1;
                const uint64_t dk3 = s == 0 ? d_k30 : d_k31;
                const int sv = V2 ? (n & 1) : 0, uv = V2 ? (n >> 1) : n;
                const uint64_t d_v = sv ? d_v1 : d_v0;
                wait_bar(&bars[(PRE ? B_RAW_FULL0 : B_QK_FULL0) + s], u & 1);
                wait_bar(&bars[B_P_TEMPTY], (n & 1) ^ 1);
                tc_fence_after();
                TRACE(2, n, 0);
                // (0) P = [q~;k~] k~^T
#pragma unroll
                for (int ks = 0; ks < K / 16; ++ks) {
                    const uint64_t off = (uint64_t)(((ks >> 2) * cfg::QK_BLK + (ks & 3) * 32) >> 4);
                    mma_ss(tmem + COL_P, dq + off, dk + off, id_p, ks > 0);
                }
                mma_commit(&bars[B_P_FULL]);
                wait_bar(&bars[B_O_EMPTY], (n & 1) ^ 1);
                // (1) OT = SA q~^T   (A from TMEM), two k-steps per 32-column block of the state as the blocks become ready
#pragma unroll
                for (int cb = 0; cb < NB; ++cb) {
                    wait_bar(&bars[B_SA_BLK0 + cb], n & 1);
                    tc_fence_after();
                    if (cb == 0) TRACE(2, n, 1);
#pragma unroll
                    for (int ks = 2 * cb; ks < 2 * cb + 2; ++ks) {
                        const uint64_t off = (uint64_t)(((ks >> 2) * cfg::QK_BLK + (ks & 3) * 32) >> 4);
                        mma_ts(tmem + COL_OT, tmem + COL_SA + ks * 8, dq + off, id_p, ks > 0);
                    }
                }
                wait_bar(&bars[B_PS_FULL], n & 1);
                wait_bar(&bars[B_V_FULL0 + sv], uv & 1);
                tc_fence_after();
                TRACE(2, n, 2);
                // (2) OT += v^T P^T
#pragma unroll
                for (int ks = 0; ks < C / 16; ++ks)
                    mma_ss(tmem + COL_OT, d_v + (uint64_t)(ks * 128), d_p + (uint64_t)(ks * 2), id_o, 1);
                mma_commit(&bars[B_O_FULL]);
                mma_commit(&bars[B_PS_EMPTY]);
                // (3) ST += v^T k~
#pragma unroll
                for (int ks = 0; ks < C / 16; ++ks)
                    mma_ss(tmem + COL_ST, d_v + (uint64_t)(ks * 128), dk3 + (uint64_t)(ks * 128), id_s, 1);
                mma_commit(&bars[B_ST_FULL]);
                mma_commit(&bars[B_QK_EMPTY0 + s]);
                mma_commit(&bars[B_V_EMPTY0 + sv]);
                TRACE(2, n, 3);
            }
            }
        }
        __syncwarp();
    } else if (warp == 16 || warp == 17) {
        // ====================== warps 16,17: causal mask of P (rows t = TMEM lanes 0..63) ======================
        const int qd = warp - 16, r = qd * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
        uint8_t *p_tile = smem + (PAIR && crank == 1 ? OFF_P1 : cfg::OFF_P);
        if (PAIR) {
            // this CTA's items m = crank, crank + 2, ...: mask P_m and deliver it to both CTAs of the pair
            const uint32_t peer = crank ^ 1u;
            const uint32_t p_remote = mapa_shared(smem_u32(p_tile), peer);
            const uint32_t bar_remote = mapa_shared(smem_u32(&bars[B_PS_FULL + crank]), peer);
            for (int m = (int)crank; m < n_items; m += 2) {
                const int j = m >> 1;
                wait_bar(&bars[B_P_FULL], j & 1);
                tc_fence_after();
                if (r == 0) TRACE(3, m, 0);
                uint32_t pr[2][32];
                tmem_ld32(tmem + lane_addr + COL_P, pr[0]);
                tmem_ld32(tmem + lane_addr + COL_P + 32, pr[1]);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&bars[B_P_TEMPTY]);
                wait_bar(&bars[B_PS_EMPTY + crank], (j & 1) ^ 1);      // (2) of item m - 2 has read this tile in BOTH CTAs
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    uint32_t w[4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int s0 = g * 8 + 2 * jj;
                        const float a = s0 <= r ? __uint_as_float(pr[s0 >> 5][s0 & 31]) : 0.f;
                        const float b = s0 + 1 <= r ? __uint_as_float(pr[(s0 + 1) >> 5][(s0 + 1) & 31]) : 0.f;
                        w[jj] = pack_bf16(a, b);
                    }
                    *reinterpret_cast<uint4 *>(p_tile + sw128_off(r, g)) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                fence_proxy_async_smem();
                named_sync(2, 64);                                     // all 64 rows written and fenced
                if (r == 0) dsmem_bulk_copy(p_remote, smem_u32(p_tile), PT_BYTES, bar_remote);
                mbar_arrive(&bars[B_PS_FULL + crank]);
                if (r == 0) TRACE(3, m, 1);
            }
        } else
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
            // row t = r ; keep s <= t ; bf16 ; P tile [t][s], 128B-swizzled
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
                *reinterpret_cast<uint4 *>(p_tile + sw128_off(r, g)) = make_uint4(w[0], w[1], w[2], w[3]);
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
            const int t0 = (n_begin + n) * C;
            wait_bar(&bars[B_O_FULL], n & 1);
            tc_fence_after();
            if (r == 0) TRACE(4, n, 0);
            uint32_t orr[2][32];
            tmem_ld32(tmem + lane_addr + COL_OT, orr[0]);
            tmem_ld32(tmem + lane_addr + COL_OT + 32, orr[1]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars[B_O_EMPTY]);
            const int nrow = min(C, T - t0);
            if (OUT32) {
                float *ob = reinterpret_cast<float *>(o) + obase + (size_t)t0 * o_tstride + v0 + r;
#pragma unroll
                for (int t = 0; t < C; ++t) {
                    if (t < nrow) ob[(size_t)t * o_tstride] = __uint_as_float(orr[t >> 5][t & 31]);
                }
            } else {
            bf16 *ob = o + obase + (size_t)t0 * o_tstride + v0 + r;
            if (nrow == C) {
#pragma unroll
                for (int t = 0; t < C; ++t) { *ob = __float2bfloat16_rn(__uint_as_float(orr[t >> 5][t & 31])); ob += o_tstride; }
            } else {
#pragma unroll
                for (int t = 0; t < C; ++t) {
                    if (t < nrow) ob[(size_t)t * o_tstride] = __float2bfloat16_rn(__uint_as_float(orr[t >> 5][t & 31]));
                }
            }
            }
            if (r == 0) TRACE(4, n, 1);
        }
    } else if (sg >= 0) {
        // ====================== state pass: warps 12-15 (+ 0-3, 4-7 with STATE2), 32-column block b by group b % NG ======================
        const int qd = warp & 3, r = qd * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
        const size_t sbase = (size_t)bh * K * V + v0 + r;          // + kappa * V
        float *hx = PAIR && piece != 0 ? sp.hx + (size_t)hx_slot * K * BV + r : nullptr;      // + kappa * BV
        if (PAIR && piece == 2) {
            // second part of a cut tile: the first part (dispatched U work-list entries earlier) has published its state
            const long long t0 = clock64();
            while (ld_acquire_gpu(sp.flags + hx_slot) == 0u) {
                if (clock64() - t0 > 4000000000LL) { printf("gla_chunk_sm100: state hand-off timeout (block %d)\n", blockIdx.x); __trap(); }
            }
        }
        // initial state -> ST (fp32) and SA (bf16)
#pragma unroll 1
        for (int cb = sg; cb < NB; cb += NG) {
            uint32_t f[32], pk[16];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float x;
                if (PAIR && piece == 2) x = hx[(size_t)(cb * 32 + j) * BV];
                else x = h0 != nullptr ? load_dyn(h0, h0_dtype, sbase + (size_t)(cb * 32 + j) * V) : 0.f;
                f[j] = __float_as_uint(x);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(__uint_as_float(f[2 * j]), __uint_as_float(f[2 * j + 1]));
            tmem_st32(tmem + lane_addr + COL_ST + cb * 32, f);
            tmem_st16(tmem + lane_addr + COL_SA + cb * 16, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        for (int cb = sg; cb < NB; cb += NG) mbar_arrive(&bars[B_SA_BLK0 + cb]);
        for (int n = 0; n < n_items; ++n) {
            const int s = n % NS, u = n / NS;
            wait_bar(&bars[(PRE ? B_RAW_FULL0 : B_QK_FULL0) + s], u & 1);   // dvec of this item is published with it
            wait_bar(&bars[B_ST_FULL], n & 1);
            tc_fence_after();
            if (r == 0 && sg == 0) TRACE(5, n, 0);
            const float *dv = dvec_slot(n);
            const bool last = n == n_items - 1;
            float rd = 1.f;
            if (ROW) rd = decay[((size_t)bh * n_total + n_begin + n) * V + v0 + r];
#pragma unroll 1
            for (int cb = sg; cb < NB; cb += NG) {
                uint32_t f[32], pk[16];
                tmem_ld32(tmem + lane_addr + COL_ST + cb * 32, f);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 d4 = ROW ? make_float4(rd, rd, rd, rd) : *reinterpret_cast<const float4 *>(dv + cb * 32 + j);
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
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(&bars[B_SA_BLK0 + cb]);        // this block of SA / ST is final for the next item
                } else if (PAIR && piece == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) hx[(size_t)(cb * 32 + j) * BV] = __uint_as_float(f[j]);
                } else if (ht != nullptr) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) ht[sbase + (size_t)(cb * 32 + j) * V] = __uint_as_float(f[j]);
                }
            }
            if (r == 0 && sg == 0) TRACE(5, n, 1);
            if (r == 0 && sg == 1) TRACE(5, n, 2);
            if (r == 0 && sg == 2) TRACE(5, n, 3);
        }
        if (PAIR && piece == 1) __threadfence();
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR && piece == 1 && tid == 0) st_release_gpu(sp.flags + hx_slot, 1u);
    if (PAIR) cluster_sync_all();        // no CTA leaves while the peer may still signal its barriers
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// How many tiles (CTA pairs) of the PAIR variant are cut in two, and the workspace that takes: with `slots` = SMs / 2 pairs
// resident at a time, U tiles run in U / slots full waves plus a last wave of U % slots; when that last wave is at most half
// full (and there is a full wave before it, and T has at least 4 chunks), its tiles are split and the halves fill it.
struct SplitPlan { int rem; size_t flag_bytes, bytes; };
static SplitPlan split_plan(int B, int H, int T, int V) {
    SplitPlan p = {0, 0, 0};
    static thread_local int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    }
    const int slots = sms / 2, U = B * H * (V / BV / 2), n_total = (T + C - 1) / C;
    if (slots <= 0 || (V / BV) % 2 != 0) return p;
    const int rem = U % slots;
    if (U < slots || rem == 0 || 2 * rem > slots || n_total < 4) return p;
    p.rem = rem;
    p.flag_bytes = (((size_t)rem * 2 * sizeof(uint32_t)) + 255) / 256 * 256;
    p.bytes = p.flag_bytes + (size_t)rem * 2 * 256 * BV * sizeof(float);       // K <= 256
    return p;
}

template <int K, int OPT = 0>
int launch(const void *q, const void *k, const void *v, const void *gk, const void *h0, int h0_dtype, void *o,
           float *ht, int B, int H, int T, int V, int bthd, float scale, cudaStream_t st, long long *trace = nullptr,
           const float *decay = nullptr, int ldq = 0, int ldk = 0, int ldv = 0, void *ws = nullptr, size_t ws_bytes = 0) {
    using cfg = Cfg<K>;
    static thread_local uint64_t configured = 0;
    if (lina_first_use_on_device(&configured))
        LINA_CUDA_OK(cudaFuncSetAttribute(gla_chunk_fwd_sm100_kernel<K, OPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)cfg::SMEM));
    // 4-D views (innermost first) (D, T, H, B) of [B,H,T,D] (bthd == 0) or [B,T,H,D] (bthd == 1) tensors;
    // 64 x 64 boxes over (D, T); rows past T read as zero.
    TMaps tm;
    {
        const uint64_t kd = K, vd = V, t = T, h = H, b = B;
        const uint64_t dims[4] = {kd, t, h, b}, vdims[4] = {vd, t, h, b};
        const uint64_t str_bhtd[4] = {2, kd * 2, t * kd * 2, h * t * kd * 2}, str_bthd[4] = {2, h * kd * 2, kd * 2, t * h * kd * 2};
        const uint64_t vstr_bhtd[4] = {2, vd * 2, t * vd * 2, h * t * vd * 2}, vstr_bthd[4] = {2, h * vd * 2, vd * 2, t * h * vd * 2};
        const uint32_t box[4] = {64, 64, 1, 1};
        const uint64_t *sk = bthd ? str_bthd : str_bhtd, *sv = bthd ? vstr_bthd : vstr_bhtd;
        int rc;
        if (ldq || ldk || ldv) {
            // operands that are column slices of wider tensors (row stride ld elements per (t, h)): the backward reads
            // V pieces of do / v in place
            const uint64_t lq = ldq ? ldq : kd, lk = ldk ? ldk : kd, lv = ldv ? ldv : vd;
            auto strides = [&](uint64_t l, uint64_t *o4) {
                if (bthd) { o4[0] = 2; o4[1] = h * l * 2; o4[2] = l * 2; o4[3] = t * h * l * 2; }
                else { o4[0] = 2; o4[1] = l * 2; o4[2] = t * l * 2; o4[3] = h * t * l * 2; }
            };
            uint64_t sq[4], sk2[4], sv2[4];
            strides(lq, sq); strides(lk, sk2); strides(lv, sv2);
            if ((rc = lina_make_tmap_bf16(&tm.q, q, 4, dims, sq, box))) return rc;
            if ((rc = lina_make_tmap_bf16(&tm.k, k, 4, dims, sk2, box))) return rc;
            if ((rc = lina_make_tmap_bf16(&tm.g, q, 4, dims, sq, box))) return rc;
            if ((rc = lina_make_tmap_bf16(&tm.v, v, 4, vdims, sv2, box))) return rc;
        } else {
        if ((rc = lina_make_tmap_bf16(&tm.q, q, 4, dims, sk, box))) return rc;
        if ((rc = lina_make_tmap_bf16(&tm.k, k, 4, dims, sk, box))) return rc;
        if ((rc = lina_make_tmap_bf16(&tm.g, gk, 4, dims, sk, box))) return rc;
        if ((rc = lina_make_tmap_bf16(&tm.v, v, 4, vdims, sv, box))) return rc;
        }
    }
    dim3 grid(V / BV, B * H);
    SplitArgs sp = {nullptr, nullptr, 0, 0, 0, 0};
    if ((OPT & 256) != 0) {
        // work list of CTA pairs (see SplitArgs): full tiles, plus `rem` tiles cut in two along T when that fills the last wave
        const int n_total = (T + C - 1) / C;
        sp.nvp = V / BV / 2;
        sp.U = B * H * sp.nvp;
        const SplitPlan pl = split_plan(B, H, T, V);
        if (ws != nullptr && ws_bytes >= pl.bytes && pl.rem > 0) {
            sp.rem = pl.rem;
            sp.n_split = (n_total + 1) / 2;
            sp.flags = reinterpret_cast<uint32_t *>(ws);
            sp.hx = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(ws) + pl.flag_bytes);
            LINA_CUDA_OK(cudaMemsetAsync(sp.flags, 0, pl.flag_bytes, st));
        }
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(2 * (sp.U + sp.rem)); lc.blockDim = dim3(NTHREADS); lc.dynamicSmemBytes = cfg::SMEM; lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        bf16 *op = (bf16 *)o;
        LINA_CUDA_OK(cudaLaunchKernelEx(&lc, gla_chunk_fwd_sm100_kernel<K, OPT>, tm, h0, h0_dtype, op, ht, T, V, H, bthd, scale,
                                        trace, decay, sp));
        return LINA_OK;
    }
    gla_chunk_fwd_sm100_kernel<K, OPT><<<grid, NTHREADS, cfg::SMEM, st>>>(tm, h0, h0_dtype, (bf16 *)o, ht, T, V, H, bthd, scale,
                                                                        trace, decay, sp);
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

static int chunk_fwd_tc(const void *q, const void *k, const void *v, const void *gk, const void *h0, int h0_dtype,
                        void *o, float *ht, int B, int H, int T, int K, int V, int bthd, float scale, void *stream) {
    LINA_REQUIRE(q && k && v && gk && o, LINA_ERR_BAD_ARG, "gla_chunk_fwd: null tensor pointer");
    LINA_REQUIRE(h0 == nullptr || lina_dtype_ok(h0_dtype), LINA_ERR_BAD_ARG, "gla_chunk_fwd: bad h0 dtype");
    cudaStream_t st = (cudaStream_t)stream;
    if (K == 64) return launch<64>(q, k, v, gk, h0, h0_dtype, o, ht, B, H, T, V, bthd, scale, st);
    if (K == 128) return launch<128>(q, k, v, gk, h0, h0_dtype, o, ht, B, H, T, V, bthd, scale, st);
    return launch<256, 0>(q, k, v, gk, h0, h0_dtype, o, ht, B, H, T, V, bthd, scale, st);
}

extern "C" int lina_gla_chunk_fwd(const void *q, const void *k, const void *v, const void *gk, const void *h0,
                                  int h0_dtype, void *o, float *ht, void *ws, int B, int H, int T, int K, int V,
                                  int dtype, float scale, void *stream) {
    (void)ws;
    if (!tc_eligible(B, H, T, K, V, dtype))
        return lina_gla_recurrent_fwd_impl(q, k, v, gk, h0, h0_dtype, o, ht, B, H, T, K, V, dtype, scale, stream);
    return chunk_fwd_tc(q, k, v, gk, h0, h0_dtype, o, ht, B, H, T, K, V, 0, scale, stream);
}

extern "C" int lina_gla_chunk_fwd_bthd(const void *q, const void *k, const void *v, const void *gk, const void *h0,
                                       int h0_dtype, void *o, float *ht, void *ws, int B, int H, int T, int K,
                                       int V, int dtype, float scale, void *stream) {
    (void)ws;
    LINA_REQUIRE(tc_eligible(B, H, T, K, V, dtype), LINA_ERR_UNSUPPORTED,
                 "gla_chunk_fwd_bthd: only the tensor-core envelope (bf16, K in {64,128,256}, V %% 128 == 0, T >= 32) "
                 "reads the [B,T,H,D] layout in place; make the tensors [B,H,T,D]-contiguous and call lina_gla_chunk_fwd");
    return chunk_fwd_tc(q, k, v, gk, h0, h0_dtype, o, ht, B, H, T, K, V, 1, scale, stream);
}

// Pre-gated operands (see OPT bit 2 of the kernel): qg = scale q e^G, kg = k e^-G [B,T,H,K] bf16, decay [B,H,NT,K] fp32
// with NT = ceil(T / 64), as written by lina_gla_prefill_prep_gated; v, o [B,T,H,V]; h0 / ht [B,H,K,V].
// Default kernel: CTA pairs sharing the score MMA (PAIR) when V / 128 is even and K >= 128; with a workspace of
// lina_gla_chunk_fwd_pregated_ws_bytes() the tiles of a less-than-half-full last wave are cut in two along T.
// A/B (lina_debug_set_variant): key 9 = 1 selects the round-1 one-CTA-per-tile kernel (2 = pairs without the T cut), key 4 = 1 the
// one-state-warpgroup form of the round-1 kernel.
static int pregated_bthd(const void *qg, const void *kg, const void *v, const float *decay, const void *h0, int h0_dtype, void *o,
                         float *ht, void *ws, size_t ws_bytes, int B, int H, int T, int K, int V, void *stream) {
    LINA_REQUIRE(qg && kg && v && decay && o, LINA_ERR_BAD_ARG, "gla_chunk_fwd_pregated: null tensor pointer");
    LINA_REQUIRE(h0 == nullptr || lina_dtype_ok(h0_dtype), LINA_ERR_BAD_ARG, "gla_chunk_fwd_pregated: bad h0 dtype");
    LINA_REQUIRE(tc_eligible(B, H, T, K, V, LINA_BF16), LINA_ERR_UNSUPPORTED,
                 "gla_chunk_fwd_pregated: outside the tensor-core envelope (K in {64,128,256}, V %% 128 == 0, T >= 32)");
    LINA_REQUIRE(((uintptr_t)decay & 15u) == 0, LINA_ERR_UNSUPPORTED, "gla_chunk_fwd_pregated: decay must be 16-byte aligned");
    LINA_REQUIRE(ws == nullptr || ((uintptr_t)ws & 255u) == 0, LINA_ERR_BAD_ARG, "gla_chunk_fwd_pregated: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int nvs = V / BV;
    if (g_lina_variant[4] == 0 && g_lina_variant[9] != 1 && nvs % 2 == 0 && K >= 128) {
        if (g_lina_variant[9] == 2) ws = nullptr;                // A/B: pairs without the T cut
        if (K == 128) return launch<128, 292>(qg, kg, v, qg, h0, h0_dtype, o, ht, B, H, T, V, 1, 1.f, st, nullptr, decay, 0, 0, 0, ws, ws_bytes);
        return launch<256, 292>(qg, kg, v, qg, h0, h0_dtype, o, ht, B, H, T, V, 1, 1.f, st, nullptr, decay, 0, 0, 0, ws, ws_bytes);
    }
    if (g_lina_variant[4] == 0) {            // three warpgroups share the state pass (variant 4 = 1: one warpgroup)
        if (K == 64) return launch<64, 36>(qg, kg, v, qg, h0, h0_dtype, o, ht, B, H, T, V, 1, 1.f, st, nullptr, decay);
        if (K == 128) return launch<128, 36>(qg, kg, v, qg, h0, h0_dtype, o, ht, B, H, T, V, 1, 1.f, st, nullptr, decay);
        return launch<256, 36>(qg, kg, v, qg, h0, h0_dtype, o, ht, B, H, T, V, 1, 1.f, st, nullptr, decay);
    }
    if (K == 64) return launch<64, 4>(qg, kg, v, qg, h0, h0_dtype, o, ht, B, H, T, V, 1, 1.f, st, nullptr, decay);
    if (K == 128) return launch<128, 4>(qg, kg, v, qg, h0, h0_dtype, o, ht, B, H, T, V, 1, 1.f, st, nullptr, decay);
    return launch<256, 4>(qg, kg, v, qg, h0, h0_dtype, o, ht, B, H, T, V, 1, 1.f, st, nullptr, decay);
}

extern "C" int lina_gla_chunk_fwd_pregated_bthd(const void *qg, const void *kg, const void *v, const float *decay,
                                                const void *h0, int h0_dtype, void *o, float *ht, int B, int H, int T,
                                                int K, int V, void *stream) {
    return pregated_bthd(qg, kg, v, decay, h0, h0_dtype, o, ht, nullptr, 0, B, H, T, K, V, stream);
}

extern "C" size_t lina_gla_chunk_fwd_pregated_ws_bytes(int B, int H, int T, int K, int V) {
    if (!tc_eligible(B, H, T, K, V, LINA_BF16) || K < 128) return 0;
    return split_plan(B, H, T, V).bytes;
}

extern "C" int lina_gla_chunk_fwd_pregated_bthd_ws(const void *qg, const void *kg, const void *v, const float *decay,
                                                   const void *h0, int h0_dtype, void *o, float *ht, void *ws, size_t ws_bytes,
                                                   int B, int H, int T, int K, int V, void *stream) {
    return pregated_bthd(qg, kg, v, decay, h0, h0_dtype, o, ht, ws, ws_bytes, B, H, T, K, V, stream);
}

// General pre-gated entry: layout [B,H,T,D] (bthd = 0) or [B,T,H,D] (bthd = 1); row_decay != 0: decay is [B,H,NT,V] and scales
// the value dim (the state's rows), out_f32 != 0: o is fp32.  Used by the backward (lina_speech_b200/fla_api/ops.py), which
// is five runs of this kernel on role-swapped / time-reversed operands.
extern "C" int lina_gla_chunk_fwd_pregated(const void *qg, const void *kg, const void *v, const float *decay,
                                           const void *h0, int h0_dtype, void *o, float *ht, int B, int H, int T, int K,
                                           int V, int bthd, int row_decay, int out_f32, int ldq, int ldk, int ldv,
                                           void *stream) {
    LINA_REQUIRE(qg && kg && v && decay && o, LINA_ERR_BAD_ARG, "gla_chunk_fwd_pregated: null tensor pointer");
    LINA_REQUIRE(h0 == nullptr || lina_dtype_ok(h0_dtype), LINA_ERR_BAD_ARG, "gla_chunk_fwd_pregated: bad h0 dtype");
    LINA_REQUIRE(tc_eligible(B, H, T, K, V, LINA_BF16), LINA_ERR_UNSUPPORTED,
                 "gla_chunk_fwd_pregated: outside the tensor-core envelope (K in {64,128,256}, V %% 128 == 0, T >= 32)");
    LINA_REQUIRE(((uintptr_t)decay & 15u) == 0, LINA_ERR_UNSUPPORTED, "gla_chunk_fwd_pregated: decay must be 16-byte aligned");
    LINA_REQUIRE((row_decay != 0) == (out_f32 != 0), LINA_ERR_UNSUPPORTED,
                 "gla_chunk_fwd_pregated: only (column decay, bf16 out) and (row decay, fp32 out) are instantiated");
    LINA_REQUIRE(ldq >= 0 && ldk >= 0 && ldv >= 0 && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 &&
                     (ldq == 0 || ldq >= K) && (ldk == 0 || ldk >= K) && (ldv == 0 || ldv >= V),
                 LINA_ERR_BAD_ARG, "gla_chunk_fwd_pregated: row strides must be 0 (dense) or multiples of 8 >= the width");
    cudaStream_t st = (cudaStream_t)stream;
    bthd = bthd ? 1 : 0;
    if (row_decay) {
        if (K == 64) return launch<64, 60>(qg, kg, v, qg, h0, h0_dtype, o, ht, B, H, T, V, bthd, 1.f, st, nullptr, decay, ldq, ldk, ldv);
        if (K == 128) return launch<128, 60>(qg, kg, v, qg, h0, h0_dtype, o, ht, B, H, T, V, bthd, 1.f, st, nullptr, decay, ldq, ldk, ldv);
        return launch<256, 60>(qg, kg, v, qg, h0, h0_dtype, o, ht, B, H, T, V, bthd, 1.f, st, nullptr, decay, ldq, ldk, ldv);
    }
    if (K == 64) return launch<64, 4>(qg, kg, v, qg, h0, h0_dtype, o, ht, B, H, T, V, bthd, 1.f, st, nullptr, decay, ldq, ldk, ldv);
    if (K == 128) return launch<128, 4>(qg, kg, v, qg, h0, h0_dtype, o, ht, B, H, T, V, bthd, 1.f, st, nullptr, decay, ldq, ldk, ldv);
    return launch<256, 4>(qg, kg, v, qg, h0, h0_dtype, o, ht, B, H, T, V, bthd, 1.f, st, nullptr, decay, ldq, ldk, ldv);
}

#ifdef LINA_GLA_TRACE
// bring-up: same kernel with a clock64 timeline of CTA (0,0): trace[6 roles][64 items][4 events]
extern "C" int lina_debug_gla_chunk_trace(const void *q, const void *k, const void *v, const void *gk, void *o, int B,
                                          int H, int T, int K, int V, float scale, long long *trace, void *stream) {
    LINA_REQUIRE(tc_eligible(B, H, T, K, V, LINA_BF16) && K == 256, LINA_ERR_UNSUPPORTED, "trace: K=256 bf16 only");
    return launch<256>(q, k, v, gk, nullptr, 0, o, nullptr, B, H, T, V, 0, scale, (cudaStream_t)stream, trace);
}

// same for the pre-gated variant (two state warpgroups): operands as for lina_gla_chunk_fwd_pregated_bthd
extern "C" int lina_debug_gla_pregated_trace(const void *qg, const void *kg, const void *v, const float *decay, void *o, int B,
                                             int H, int T, int K, int V, long long *trace, void *stream) {
    LINA_REQUIRE(tc_eligible(B, H, T, K, V, LINA_BF16) && K == 256, LINA_ERR_UNSUPPORTED, "trace: K=256 bf16 only");
    if (g_lina_variant[9] != 1 && (V / BV) % 2 == 0)       // the CTA-pair kernel (rank 0 of the first pair is traced)
        return launch<256, 292>(qg, kg, v, qg, nullptr, 0, o, nullptr, B, H, T, V, 1, 1.f, (cudaStream_t)stream, trace, decay);
    return launch<256, 36>(qg, kg, v, qg, nullptr, 0, o, nullptr, B, H, T, V, 1, 1.f, (cudaStream_t)stream, trace, decay);
}
#endif  // LINA_GLA_TRACE
