// Element-wise passes around the tensor-core GLA BACKWARD (bf16, chunk C = 64).  The backward itself is five runs of the
// pre-gated forward kernel (gla_chunk_sm100.cu, see lina_speech_b200/fla_api/ops.py:_bwd_tc for the identities, which
// regroup FLA/fla/ops/gla/chunk.py:140-341 and FLA/fla/ops/common/chunk_h.py:111-189); these kernels build its operands
// and finish its outputs, each in ONE pass over the tensors (the first version did this with ~25 torch ops per layer:
// 65 ms of a 300 ms training step at bs8 x seq4096).
//
//   prep   : q, k, gk -> k~ = k e^-G (forward time), q^ = scale q e^{G-G_C} and k^ = k e^{G_C-G} (time-REVERSED, padded to
//            whole chunks), D = e^{G_C} in forward and reversed chunk order            (G = in-chunk cumsum of gk, fp32)
//   flip2  : do, v -> time-reversed, chunk-padded copies
//   post   : dq = (sum of dq~ pieces) * scale e^G ; dk = (sum of reversed dk^ pieces) * e^{G_C-G} ;
//            dgk_local = in-chunk reversed cumsum of (dq q - dk k), chunk totals
//   finish : dgk = dgk_local + carry[chunk]      (carry = totals of the later chunks [+ the dht term], a tiny host-side scan)
// One thread = 4 channels x one chunk (64 rows, serial: both cumsums are thread-local); 8-byte loads, packed fp32x2 math.
#include "common.cuh"
#include "sm100.cuh"
#include "packed.cuh"

namespace {

using sm100::ex2_approx;
using sm100::pack_bf16;

constexpr int C = 64;
constexpr int RB = 8;                                   // rows per load batch
constexpr float LOG2E = 1.44269504088896340736f;

// element strides (batch, head, time) of a [B,H,T,K] (bthd = 0) or [B,T,H,K] (bthd = 1) tensor; channels are contiguous
struct Lay { size_t sb, sh, st; };
Lay make_lay(int bthd, int H, int Tn, int K) {
    Lay l;
    if (bthd) { l.sb = (size_t)Tn * H * K; l.sh = (size_t)K; l.st = (size_t)H * K; }
    else { l.sb = (size_t)H * Tn * K; l.sh = (size_t)Tn * K; l.st = (size_t)K; }
    return l;
}

struct ChunkIdx { int cg, n; long long bh; bool ok; };
__device__ __forceinline__ ChunkIdx chunk_index(int ng, int NT, long long BH) {
    const long long idx = (long long)blockIdx.x * 128 + threadIdx.x;
    ChunkIdx c;
    c.cg = (int)(idx % ng);
    const long long rest = idx / ng;
    c.n = (int)(rest % NT);
    c.bh = rest / NT;
    c.ok = c.bh < BH;
    return c;
}

// sum of gk over the valid rows of the chunk, in log2 units
__device__ __forceinline__ void chunk_gate_total(const bf16 *g, size_t K, int nrow, float2 (&GC)[2]) {
    GC[0] = GC[1] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int r0 = 0; r0 < nrow; r0 += RB) {
        uint2 rg[RB];
#pragma unroll
        for (int i = 0; i < RB; ++i)
            rg[i] = r0 + i < nrow ? *reinterpret_cast<const uint2 *>(g + (size_t)(r0 + i) * K) : make_uint2(0, 0);
#pragma unroll
        for (int i = 0; i < RB; ++i) {
            GC[0] = __fadd2_rn(GC[0], bf2_to_f2(rg[i].x));
            GC[1] = __fadd2_rn(GC[1], bf2_to_f2(rg[i].y));
        }
    }
    GC[0] = __fmul2_rn(GC[0], make_float2(LOG2E, LOG2E));
    GC[1] = __fmul2_rn(GC[1], make_float2(LOG2E, LOG2E));
}

__global__ void __launch_bounds__(128, 4)
gla_bwd_prep_kernel(const bf16 *__restrict__ q, const bf16 *__restrict__ k, const bf16 *__restrict__ gk,
                    bf16 *__restrict__ kt, bf16 *__restrict__ qh_r, bf16 *__restrict__ kh_r, float *__restrict__ D,
                    float *__restrict__ Dr, long long BH, int H, int T, int K, float log2_scale, Lay lt, Lay lp) {
    const int NT = (T + C - 1) / C, Tp = NT * C;
    const ChunkIdx ci = chunk_index(K / 4, NT, BH);
    if (!ci.ok) return;
    const int d0 = ci.cg * 4, t0 = ci.n * C;
    const int nrow = min(C, T - t0);
    const size_t bb = (size_t)(ci.bh / H), hh = (size_t)(ci.bh % H);
    const size_t in0 = bb * lt.sb + hh * lt.sh + (size_t)t0 * lt.st + d0;
    const bf16 *qp = q + in0, *kp = k + in0, *gp = gk + in0;
    bf16 *ktp = kt + in0;
    const size_t rbase = bb * lp.sb + hh * lp.sh + d0;          // reversed arrays: row Tp-1-t
    float2 GC[2];
    chunk_gate_total(gp, lt.st, nrow, GC);
    const float4 dec = make_float4(ex2_approx(GC[0].x), ex2_approx(GC[0].y), ex2_approx(GC[1].x), ex2_approx(GC[1].y));
    *reinterpret_cast<float4 *>(D + ((size_t)ci.bh * NT + ci.n) * K + d0) = dec;
    *reinterpret_cast<float4 *>(Dr + ((size_t)ci.bh * NT + (NT - 1 - ci.n)) * K + d0) = dec;
    float2 G[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll 1
    for (int r0 = 0; r0 < C; r0 += RB) {
        uint2 rq[RB], rk[RB], rg[RB];
#pragma unroll
        for (int i = 0; i < RB; ++i) {
            const bool ok = r0 + i < nrow;
            const size_t off = (size_t)(r0 + i) * lt.st;
            rq[i] = ok ? *reinterpret_cast<const uint2 *>(qp + off) : make_uint2(0, 0);
            rk[i] = ok ? *reinterpret_cast<const uint2 *>(kp + off) : make_uint2(0, 0);
            rg[i] = ok ? *reinterpret_cast<const uint2 *>(gp + off) : make_uint2(0, 0);
        }
#pragma unroll
        for (int i = 0; i < RB; ++i) {
            const int t = t0 + r0 + i;
            const bool ok = r0 + i < nrow;
            uint32_t okt[2], oqh[2], okh[2];
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const float2 qv = bf2_to_f2(p == 0 ? rq[i].x : rq[i].y), kv = bf2_to_f2(p == 0 ? rk[i].x : rk[i].y);
                G[p] = __ffma2_rn(bf2_to_f2(p == 0 ? rg[i].x : rg[i].y), make_float2(LOG2E, LOG2E), G[p]);
                const float2 e_n = make_float2(ex2_approx(-G[p].x), ex2_approx(-G[p].y));                       // e^-G
                const float2 e_q = make_float2(ex2_approx(G[p].x - GC[p].x + log2_scale), ex2_approx(G[p].y - GC[p].y + log2_scale));
                const float2 e_k = make_float2(ex2_approx(GC[p].x - G[p].x), ex2_approx(GC[p].y - G[p].y));
                const float2 a = __fmul2_rn(kv, e_n), b = __fmul2_rn(qv, e_q), c = __fmul2_rn(kv, e_k);
                okt[p] = pack_bf16(a.x, a.y); oqh[p] = pack_bf16(b.x, b.y); okh[p] = pack_bf16(c.x, c.y);
            }
            // rows past the end of the sequence (ragged last chunk) are zero rows of the padded, reversed operands
            const uint2 zq = ok ? make_uint2(oqh[0], oqh[1]) : make_uint2(0, 0);
            const uint2 zk = ok ? make_uint2(okh[0], okh[1]) : make_uint2(0, 0);
            const size_t roff = rbase + (size_t)(Tp - 1 - t) * lp.st;
            *reinterpret_cast<uint2 *>(qh_r + roff) = zq;
            *reinterpret_cast<uint2 *>(kh_r + roff) = zk;
            if (ok) *reinterpret_cast<uint2 *>(ktp + (size_t)(r0 + i) * lt.st) = make_uint2(okt[0], okt[1]);
        }
    }
}

// y[bh, Tp-1-t, :] = x[bh, t, :] for t < T, zero rows for T <= t < Tp; two tensors per launch (blockIdx.y)
__global__ void __launch_bounds__(256)
time_reverse_pad_kernel(const bf16 *__restrict__ a, const bf16 *__restrict__ b, bf16 *__restrict__ ar, bf16 *__restrict__ br,
                        long long BH, int T, int Tp, int Dm) {
    const bf16 *x = blockIdx.y == 0 ? a : b;
    bf16 *y = blockIdx.y == 0 ? ar : br;
    const int nv = Dm / 8;
    const long long total = BH * Tp * nv;
#pragma unroll 1
    for (long long i0 = (long long)blockIdx.x * 1024 + threadIdx.x; i0 < total; i0 += (long long)gridDim.x * 1024) {
        uint4 val[4];
        long long dst[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long i = i0 + u * 256;
            dst[u] = -1;
            val[u] = make_uint4(0, 0, 0, 0);
            if (i < total) {
                const int dv = (int)(i % nv);
                const long long rest = i / nv;
                const int tr = (int)(rest % Tp);
                const long long bh = rest / Tp;
                const int t = Tp - 1 - tr;
                dst[u] = i;
                if (t < T) val[u] = *reinterpret_cast<const uint4 *>(x + ((size_t)bh * T + t) * Dm + (size_t)dv * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (dst[u] >= 0) *reinterpret_cast<uint4 *>(y + (size_t)dst[u] * 8) = val[u];
    }
}

constexpr int RBP = 4;                                  // rows per load batch of the post pass (5 streams per row)
__global__ void __launch_bounds__(128, 4)
gla_bwd_post_kernel(const float *__restrict__ dqa, const float *__restrict__ dqb, const float *__restrict__ dka,
                    const float *__restrict__ dkb, const bf16 *__restrict__ q, const bf16 *__restrict__ k,
                    const bf16 *__restrict__ gk, bf16 *__restrict__ dq, bf16 *__restrict__ dk,
                    float *__restrict__ dgk_local, float *__restrict__ totals, long long BH, int H, int T, int K,
                    float log2_scale, Lay lt, Lay lp) {
    const int NT = (T + C - 1) / C, Tp = NT * C;
    const ChunkIdx ci = chunk_index(K / 4, NT, BH);
    if (!ci.ok) return;
    const int d0 = ci.cg * 4, t0 = ci.n * C;
    const int nrow = min(C, T - t0);
    const size_t bb = (size_t)(ci.bh / H), hh = (size_t)(ci.bh % H);
    const size_t in0 = bb * lt.sb + hh * lt.sh + (size_t)t0 * lt.st + d0;
    const size_t rbase = bb * lp.sb + hh * lp.sh + d0;
    float2 GC[2];
    chunk_gate_total(gk + in0, lt.st, nrow, GC);
    float2 G[2] = {GC[0], GC[1]};                                // inclusive cumsum at the last valid row
    float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll 1
    for (int r1 = nrow; r1 > 0; r1 -= RBP) {                       // rows r1-1 down to r1-RBP
        uint2 rq[RBP], rk[RBP], rg[RBP];
        float4 fa[RBP], fk[RBP];
#pragma unroll
        for (int i = 0; i < RBP; ++i) {
            const int r = r1 - 1 - i;
            const bool ok = r >= 0;
            const size_t off = in0 + (size_t)(ok ? r : 0) * lt.st;
            const size_t roff = rbase + (size_t)(Tp - 1 - (t0 + (ok ? r : 0))) * lp.st;
            rq[i] = ok ? *reinterpret_cast<const uint2 *>(q + off) : make_uint2(0, 0);
            rk[i] = ok ? *reinterpret_cast<const uint2 *>(k + off) : make_uint2(0, 0);
            rg[i] = ok ? *reinterpret_cast<const uint2 *>(gk + off) : make_uint2(0, 0);
            fa[i] = ok ? *reinterpret_cast<const float4 *>(dqa + off) : make_float4(0.f, 0.f, 0.f, 0.f);
            fk[i] = ok ? *reinterpret_cast<const float4 *>(dka + roff) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok && dqb != nullptr) {
                const float4 x = *reinterpret_cast<const float4 *>(dqb + off), y = *reinterpret_cast<const float4 *>(dkb + roff);
                fa[i] = make_float4(fa[i].x + x.x, fa[i].y + x.y, fa[i].z + x.z, fa[i].w + x.w);
                fk[i] = make_float4(fk[i].x + y.x, fk[i].y + y.y, fk[i].z + y.z, fk[i].w + y.w);
            }
        }
#pragma unroll
        for (int i = 0; i < RBP; ++i) {
            const int r = r1 - 1 - i;
            if (r < 0) break;
            uint32_t oq[2], ok_[2];
            float2 lo[2];
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const float2 qv = bf2_to_f2(p == 0 ? rq[i].x : rq[i].y), kv = bf2_to_f2(p == 0 ? rk[i].x : rk[i].y);
                const float2 e_q = make_float2(ex2_approx(G[p].x + log2_scale), ex2_approx(G[p].y + log2_scale));
                const float2 e_k = make_float2(ex2_approx(GC[p].x - G[p].x), ex2_approx(GC[p].y - G[p].y));
                const float2 dqt = p == 0 ? make_float2(fa[i].x, fa[i].y) : make_float2(fa[i].z, fa[i].w);
                const float2 dkt = p == 0 ? make_float2(fk[i].x, fk[i].y) : make_float2(fk[i].z, fk[i].w);
                const float2 dqv = __fmul2_rn(dqt, e_q), dkv = __fmul2_rn(dkt, e_k);
                acc[p] = __ffma2_rn(dqv, qv, acc[p]);
                acc[p] = __ffma2_rn(__fmul2_rn(dkv, make_float2(-1.f, -1.f)), kv, acc[p]);
                lo[p] = acc[p];
                oq[p] = pack_bf16(dqv.x, dqv.y);
                ok_[p] = pack_bf16(dkv.x, dkv.y);
                G[p] = __ffma2_rn(bf2_to_f2(p == 0 ? rg[i].x : rg[i].y), make_float2(-LOG2E, -LOG2E), G[p]);   // -> G_{t-1}
            }
            const size_t off = in0 + (size_t)r * lt.st;
            *reinterpret_cast<uint2 *>(dq + off) = make_uint2(oq[0], oq[1]);
            *reinterpret_cast<uint2 *>(dk + off) = make_uint2(ok_[0], ok_[1]);
            *reinterpret_cast<float4 *>(dgk_local + off) = make_float4(lo[0].x, lo[0].y, lo[1].x, lo[1].y);
        }
    }
    *reinterpret_cast<float4 *>(totals + ((size_t)ci.bh * NT + ci.n) * K + d0) = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
}

__global__ void __launch_bounds__(256)
gla_bwd_dgk_finish_kernel(const float *__restrict__ dgk_local, const float *__restrict__ carry, bf16 *__restrict__ dgk,
                          long long BH, int H, int T, int K, Lay lt) {
    const int NT = (T + C - 1) / C, nv = K / 4;
    const long long total = BH * T * nv;
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int dv = (int)(i % nv);                                   // enumerate in MEMORY order of the layout: (.., t|h, h|t, dv)
    const long long rest = i / nv;
    long long bh;
    int t;
    if (lt.sh < lt.st) {                                            // [B,T,H,K]
        const int hh = (int)(rest % H);
        const long long r2 = rest / H;
        t = (int)(r2 % T);
        bh = (r2 / T) * H + hh;
    } else {
        t = (int)(rest % T);
        bh = rest / T;
    }
    const float4 a = *reinterpret_cast<const float4 *>(dgk_local + (size_t)i * 4);
    const float4 c = *reinterpret_cast<const float4 *>(carry + ((size_t)bh * NT + t / C) * K + (size_t)dv * 4);
    *reinterpret_cast<uint2 *>(dgk + (size_t)i * 4) = make_uint2(pack_bf16(a.x + c.x, a.y + c.y), pack_bf16(a.z + c.z, a.w + c.w));
}

bool al16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

}  // namespace

extern "C" int lina_gla_bwd_prep(const void *q, const void *k, const void *gk, void *kt, void *qh_r, void *kh_r, float *D,
                                 float *Dr, int B, int H, int T, int K, int bthd, float scale, void *stream) {
    LINA_REQUIRE(q && k && gk && kt && qh_r && kh_r && D && Dr, LINA_ERR_BAD_ARG, "gla_bwd_prep: null pointer");
    LINA_REQUIRE(B > 0 && H > 0 && T > 0 && K > 0 && K % 4 == 0 && scale > 0.f, LINA_ERR_BAD_ARG, "gla_bwd_prep: bad size");
    LINA_REQUIRE(al16(q) && al16(k) && al16(gk) && al16(kt) && al16(qh_r) && al16(kh_r) && al16(D) && al16(Dr),
                 LINA_ERR_UNSUPPORTED, "gla_bwd_prep: tensors must be 16-byte aligned");
    const int NT = (T + C - 1) / C;
    const long long n = (long long)B * H * NT * (K / 4);
    gla_bwd_prep_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        (const bf16 *)q, (const bf16 *)k, (const bf16 *)gk, (bf16 *)kt, (bf16 *)qh_r, (bf16 *)kh_r, D, Dr, (long long)B * H,
        H, T, K, log2f(scale), make_lay(bthd, H, T, K), make_lay(bthd, H, NT * C, K));
    LINA_LAUNCH_OK("gla_bwd_prep_kernel");
    return LINA_OK;
}

extern "C" int lina_time_reverse_pad2(const void *a, const void *b, void *a_r, void *b_r, long long BH, int T, int Tp, int Dm,
                                      void *stream) {
    LINA_REQUIRE(a && b && a_r && b_r && BH > 0 && T > 0 && Tp >= T && Dm > 0 && Dm % 8 == 0, LINA_ERR_BAD_ARG,
                 "time_reverse_pad2: bad argument");
    LINA_REQUIRE(al16(a) && al16(b) && al16(a_r) && al16(b_r), LINA_ERR_UNSUPPORTED, "time_reverse_pad2: 16-byte alignment");
    const long long total = BH * Tp * (Dm / 8);
    long long nblk = (total + 1023) / 1024;
    if (nblk > 148LL * 64) nblk = 148LL * 64;
    dim3 grid((unsigned)nblk, 2);
    time_reverse_pad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16 *)a, (const bf16 *)b, (bf16 *)a_r, (bf16 *)b_r,
                                                                    BH, T, Tp, Dm);
    LINA_LAUNCH_OK("time_reverse_pad_kernel");
    return LINA_OK;
}

extern "C" int lina_gla_bwd_post(const float *dqa, const float *dqb, const float *dka, const float *dkb, const void *q,
                                 const void *k, const void *gk, void *dq, void *dk, float *dgk_local, float *totals, int B,
                                 int H, int T, int K, int bthd, float scale, void *stream) {
    LINA_REQUIRE(dqa && dka && q && k && gk && dq && dk && dgk_local && totals, LINA_ERR_BAD_ARG, "gla_bwd_post: null pointer");
    LINA_REQUIRE((dqb == nullptr) == (dkb == nullptr), LINA_ERR_BAD_ARG, "gla_bwd_post: pass both second pieces or none");
    LINA_REQUIRE(B > 0 && H > 0 && T > 0 && K > 0 && K % 4 == 0 && scale > 0.f, LINA_ERR_BAD_ARG, "gla_bwd_post: bad size");
    const int NT = (T + C - 1) / C;
    const long long n = (long long)B * H * NT * (K / 4);
    gla_bwd_post_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        dqa, dqb, dka, dkb, (const bf16 *)q, (const bf16 *)k, (const bf16 *)gk, (bf16 *)dq, (bf16 *)dk, dgk_local, totals,
        (long long)B * H, H, T, K, log2f(scale), make_lay(bthd, H, T, K), make_lay(bthd, H, NT * C, K));
    LINA_LAUNCH_OK("gla_bwd_post_kernel");
    return LINA_OK;
}

extern "C" int lina_gla_bwd_dgk_finish(const float *dgk_local, const float *carry, void *dgk, int B, int H, int T, int K,
                                       int bthd, void *stream) {
    LINA_REQUIRE(dgk_local && carry && dgk && B > 0 && H > 0 && T > 0 && K > 0 && K % 4 == 0, LINA_ERR_BAD_ARG,
                 "gla_bwd_dgk_finish: bad argument");
    const long long n = (long long)B * H * T * (K / 4);
    gla_bwd_dgk_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        dgk_local, carry, (bf16 *)dgk, (long long)B * H, H, T, K, make_lay(bthd, H, T, K));
    LINA_LAUNCH_OK("gla_bwd_dgk_finish_kernel");
    return LINA_OK;
}
