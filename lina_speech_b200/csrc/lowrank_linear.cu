// Rank-R expansion linear  out[m, n] = bias[n] + sum_{r < R} x[m, r] * W[n, r]   (bf16 in / out, fp32 accumulate).
//
// This is gk_proj[1] of GatedLinearAttention (model/gla.py:96-97: nn.Linear(gate_low_rank_dim = 16, key_dim, bias=True)) on a
// whole sequence: M = B*T rows, K = 16.  As a library GEMM it is a K = 16 problem with a 134 MB output at the bench shape and
// ran at 0.31 ms per call (12 % of the forward step, 13 calls); it is an HBM-write-bound outer-product expansion, so one
// thread keeps the R weights of 4 output channels in registers (packed fp32 pairs) and walks 64 rows whose R inputs sit in
// shared memory as duplicated pairs: 2 R packed FMAs + R/2 shared loads per 4 outputs.
#include "common.cuh"
#include "packed.cuh"

namespace {

constexpr int LR_THREADS = 256;      // 256 threads x 4 channels = 1024 output channels per block column
constexpr int LR_ROWS = 64;          // rows per block

template <int R>
__global__ void __launch_bounds__(LR_THREADS)
lowrank_linear_bf16_kernel(const bf16 *__restrict__ x, long long ldx, const bf16 *__restrict__ W, const bf16 *__restrict__ bias,
                           bf16 *__restrict__ out, long long ldo, int M, int N) {
    __shared__ float2 xs[LR_ROWS][R];            // {x, x}: the packed-FMA operand
    const int tid = threadIdx.x;
    const int n0 = (blockIdx.y * LR_THREADS + tid) * 4;
    const int m0 = blockIdx.x * LR_ROWS;
    const int rows = min(LR_ROWS, M - m0);
    // stage the input rows: LR_ROWS x R bf16, 4 values per thread and pass
    for (int i = tid; i < LR_ROWS * R / 4; i += LR_THREADS) {
        const int row = i / (R / 4), part = i - row * (R / 4);
        uint2 raw = make_uint2(0u, 0u);
        if (row < rows) raw = *reinterpret_cast<const uint2 *>(x + (size_t)(m0 + row) * ldx + part * 4);
        const float2 a = bf2_to_f2(raw.x), b = bf2_to_f2(raw.y);
        xs[row][part * 4 + 0] = make_float2(a.x, a.x);
        xs[row][part * 4 + 1] = make_float2(a.y, a.y);
        xs[row][part * 4 + 2] = make_float2(b.x, b.x);
        xs[row][part * 4 + 3] = make_float2(b.y, b.y);
    }
    const bool active = n0 < N;
    float2 w[R][2];                              // (channel 0, 1), (channel 2, 3) of this thread, per rank index
    float2 b01 = make_float2(0.f, 0.f), b23 = make_float2(0.f, 0.f);
    if (active) {
        float wf[4][R];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int j = 0; j < R / 8; ++j) {
                const uint4 raw = *reinterpret_cast<const uint4 *>(W + (size_t)(n0 + c) * R + j * 8);
                const float2 p0 = bf2_to_f2(raw.x), p1 = bf2_to_f2(raw.y), p2 = bf2_to_f2(raw.z), p3 = bf2_to_f2(raw.w);
                wf[c][j * 8 + 0] = p0.x; wf[c][j * 8 + 1] = p0.y; wf[c][j * 8 + 2] = p1.x; wf[c][j * 8 + 3] = p1.y;
                wf[c][j * 8 + 4] = p2.x; wf[c][j * 8 + 5] = p2.y; wf[c][j * 8 + 6] = p3.x; wf[c][j * 8 + 7] = p3.y;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) { w[r][0] = make_float2(wf[0][r], wf[1][r]); w[r][1] = make_float2(wf[2][r], wf[3][r]); }
        if (bias != nullptr) {
            const uint2 raw = *reinterpret_cast<const uint2 *>(bias + n0);
            b01 = bf2_to_f2(raw.x); b23 = bf2_to_f2(raw.y);
        }
    }
    __syncthreads();
    if (!active) return;
    bf16 *o = out + (size_t)m0 * ldo + n0;
#pragma unroll 2
    for (int row = 0; row < rows; ++row) {
        float2 a01 = b01, a23 = b23;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float2 xv = xs[row][r];
            a01 = __ffma2_rn(w[r][0], xv, a01);
            a23 = __ffma2_rn(w[r][1], xv, a23);
        }
        uint2 pk;
        pk.x = f2_to_bf2(a01);
        pk.y = f2_to_bf2(a23);
        *reinterpret_cast<uint2 *>(o + (size_t)row * ldo) = pk;
    }
}

}  // namespace

extern "C" int lina_lowrank_linear(const void *x, long long ldx, const void *W, const void *bias, void *out, long long ldo,
                                   int M, int N, int R, int dtype, void *stream) {
    LINA_REQUIRE(x && W && out, LINA_ERR_BAD_ARG, "lowrank_linear: null pointer");
    LINA_REQUIRE(M > 0 && N > 0, LINA_ERR_BAD_ARG, "lowrank_linear: non-positive size");
    LINA_REQUIRE(dtype == LINA_BF16, LINA_ERR_UNSUPPORTED, "lowrank_linear: bf16 only (dtype %d)", dtype);
    LINA_REQUIRE(R == 8 || R == 16 || R == 32, LINA_ERR_UNSUPPORTED, "lowrank_linear: rank %d (8, 16 or 32)", R);
    LINA_REQUIRE(N % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && ldx >= R && ldo >= N, LINA_ERR_UNSUPPORTED,
                 "lowrank_linear: N, ldx, ldo must be multiples of 4 (8-byte vectors)");
    LINA_REQUIRE(((uintptr_t)x & 7u) == 0 && ((uintptr_t)out & 7u) == 0 && ((uintptr_t)W & 15u) == 0 &&
                     (bias == nullptr || ((uintptr_t)bias & 7u) == 0),
                 LINA_ERR_UNSUPPORTED, "lowrank_linear: x / out / bias must be 8-byte aligned, W 16-byte aligned");
    const long long gx = ((long long)M + LR_ROWS - 1) / LR_ROWS;
    LINA_REQUIRE(gx <= 2147483647LL, LINA_ERR_UNSUPPORTED, "lowrank_linear: grid too large");
    dim3 grid((unsigned)gx, (unsigned)((N / 4 + LR_THREADS - 1) / LR_THREADS));
    cudaStream_t st = (cudaStream_t)stream;
    const bf16 *xp = (const bf16 *)x, *wp = (const bf16 *)W, *bp = (const bf16 *)bias;
    bf16 *op = (bf16 *)out;
    if (R == 8) lowrank_linear_bf16_kernel<8><<<grid, LR_THREADS, 0, st>>>(xp, ldx, wp, bp, op, ldo, M, N);
    else if (R == 16) lowrank_linear_bf16_kernel<16><<<grid, LR_THREADS, 0, st>>>(xp, ldx, wp, bp, op, ldo, M, N);
    else lowrank_linear_bf16_kernel<32><<<grid, LR_THREADS, 0, st>>>(xp, ldx, wp, bp, op, ldo, M, N);
    LINA_LAUNCH_OK("lowrank_linear_bf16_kernel");
    return LINA_OK;
}
