// Single-CTA tcgen05 probe: D[128,N] = A[128,KD] * B[N,KD]^T in bf16 -> fp32, with every operand
// placement the GLA kernel uses (smem K-major, smem MN-major, A from TMEM).  Exists to pin the
// descriptor conventions of sm100.cuh on real hardware (tests/test_umma_probe_gpu.py).
#include "../common.cuh"
#include "../sm100.cuh"

using namespace sm100;

namespace {

__global__ void __launch_bounds__(128)
umma_probe_kernel(const float *__restrict__ A, const float *__restrict__ Bm, float *__restrict__ D, int N, int KD,
                  int a_mode, int b_mode, int swap) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t gA_k = (128 + 1) * 16, gA_mn = (KD + 1) * 16;
    const uint32_t gB_k = (N + 1) * 16, gB_mn = (KD + 1) * 16;
    uint8_t *a_tile = smem;
    uint8_t *b_tile = smem + 40 * 1024;
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    // ---- A into smem
    for (int i = tid; i < 128 * KD; i += 128) {
        const int m = i / KD, k = i % KD;
        const bf16 v = __float2bfloat16_rn(A[i]);
        if (a_mode == 0) *reinterpret_cast<bf16 *>(a_tile + (k / 8) * gA_k + m * 16 + (k % 8) * 2) = v;
        else if (a_mode == 1) *reinterpret_cast<bf16 *>(a_tile + (m / 8) * gA_mn + k * 16 + (m % 8) * 2) = v;
    }
    for (int i = tid; i < N * KD; i += 128) {
        const int n = i / KD, k = i % KD;
        const bf16 v = __float2bfloat16_rn(Bm[i]);
        if (b_mode == 0) *reinterpret_cast<bf16 *>(b_tile + (k / 8) * gB_k + n * 16 + (k % 8) * 2) = v;
        else *reinterpret_cast<bf16 *>(b_tile + (n / 8) * gB_mn + k * 16 + (n % 8) * 2) = v;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    const uint32_t t_d = tbase, t_a = tbase + 256;
    if (a_mode == 2) {            // A operand in TMEM: lane = row m, column c holds (k=2c, k=2c+1)
        const int m = tid;
        for (int c0 = 0; c0 < KD / 2; c0 += 16) {
            uint32_t r[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int k = 2 * (c0 + j);
                r[j] = pack_bf16(A[m * KD + k], A[m * KD + k + 1]);
            }
            tmem_st16(t_a + ((uint32_t)(warp * 32) << 16) + c0, r);
        }
        tmem_st_wait();
        tc_fence_before();
    }
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        const uint32_t idesc = idesc_bf16(128, N, a_mode == 1, b_mode == 1);
        for (int ks = 0; ks < KD / 16; ++ks) {
            uint32_t lbo, sbo, start;
            if (b_mode == 0) { start = smem_u32(b_tile) + ks * 2 * gB_k; lbo = gB_k; sbo = 128; }
            else { start = smem_u32(b_tile) + ks * 256; lbo = 128; sbo = gB_mn; }
            const uint64_t bd = swap ? smem_desc(start, sbo, lbo) : smem_desc(start, lbo, sbo);
            if (a_mode == 2) {
                mma_ts(t_d, t_a + ks * 8, bd, idesc, ks > 0);
            } else {
                if (a_mode == 0) { start = smem_u32(a_tile) + ks * 2 * gA_k; lbo = gA_k; sbo = 128; }
                else { start = smem_u32(a_tile) + ks * 256; lbo = 128; sbo = gA_mn; }
                const uint64_t ad = swap ? smem_desc(start, sbo, lbo) : smem_desc(start, lbo, sbo);
                mma_ss(t_d, ad, bd, idesc, ks > 0);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_d + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tbase);
}

}  // namespace

extern "C" int lina_debug_umma_probe(const float *A, const float *B, float *D, int N, int KD, int a_mode, int b_mode,
                                     int swap, void *stream) {
    LINA_REQUIRE(A && B && D, LINA_ERR_BAD_ARG, "umma_probe: null pointer");
    LINA_REQUIRE(N % 32 == 0 && N >= 32 && N <= 256 && KD % 32 == 0 && KD >= 32 && KD <= 128, LINA_ERR_BAD_ARG,
                 "umma_probe: need N in [32,256] %% 32, KD in [32,128] %% 32");
    const int smem = 120 * 1024;
    LINA_CUDA_OK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, KD, a_mode, b_mode, swap);
    LINA_LAUNCH_OK("umma_probe_kernel");
    return LINA_OK;
}

// ---- M = 64 layout probe (round-2 bring-up, NOT yet run on hardware): same no-swizzle K-major operands, instruction shape
// M x N x 16 with M in {64, 128}; dumps the RAW accumulator tile -- all 128 TMEM lanes x N columns -- so that the lane mapping
// of an M = 64 accumulator can be read off (profiles/probe_m64.py).  Motivation: the score MMA of the GLA kernel needs only
// its 64 q~ rows; at M = 128 half of its shared-memory A reads (the kernel's binding resource) are a by-product.
namespace {

__global__ void __launch_bounds__(128)
umma_probe_m_kernel(const float *__restrict__ A, const float *__restrict__ Bm, float *__restrict__ D, int M, int N, int KD) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t gA_k = (M + 1) * 16, gB_k = (N + 1) * 16;
    uint8_t *a_tile = smem, *b_tile = smem + 40 * 1024;
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    for (int i = tid; i < M * KD; i += 128) {
        const int m = i / KD, k = i % KD;
        *reinterpret_cast<bf16 *>(a_tile + (k / 8) * gA_k + m * 16 + (k % 8) * 2) = __float2bfloat16_rn(A[i]);
    }
    for (int i = tid; i < N * KD; i += 128) {
        const int n = i / KD, k = i % KD;
        *reinterpret_cast<bf16 *>(b_tile + (k / 8) * gB_k + n * 16 + (k % 8) * 2) = __float2bfloat16_rn(Bm[i]);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    {   // poison the tile so untouched cells are recognisable
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(-12345.f);
        for (int c0 = 0; c0 < N; c0 += 32) tmem_st32(tbase + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_st_wait();
        tc_fence_before();
    }
    __syncthreads();
    tc_fence_after();
    if (warp == 0 && elect_one_sync()) {
        const uint32_t idesc = idesc_bf16(M, N, 0, 0);
        for (int ks = 0; ks < KD / 16; ++ks)
            mma_ss(tbase, smem_desc(smem_u32(a_tile) + ks * 2 * gA_k, gA_k, 128), smem_desc(smem_u32(b_tile) + ks * 2 * gB_k, gB_k, 128),
                   idesc, ks > 0);
        mma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tbase);
}

}  // namespace

extern "C" int lina_debug_umma_probe_m(const float *A, const float *B, float *D, int M, int N, int KD, void *stream) {
    LINA_REQUIRE(A && B && D, LINA_ERR_BAD_ARG, "umma_probe_m: null pointer");
    LINA_REQUIRE((M == 64 || M == 128) && N % 32 == 0 && N >= 32 && N <= 256 && KD % 16 == 0 && KD >= 16 && KD <= 128,
                 LINA_ERR_BAD_ARG, "umma_probe_m: need M in {64,128}, N in [32,256] %% 32, KD in [16,128] %% 16");
    const int smem = 120 * 1024;
    LINA_CUDA_OK(cudaFuncSetAttribute(umma_probe_m_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_probe_m_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, M, N, KD);
    LINA_LAUNCH_OK("umma_probe_m_kernel");
    return LINA_OK;
}

// ---- second probe: 128-byte-swizzled operands (K-major / MN-major) and a TMA-loaded A ----------------------
#include "../tma.cuh"

namespace {

__global__ void __launch_bounds__(128)
umma_probe_sw128_kernel(const float *__restrict__ A, const float *__restrict__ Bm, float *__restrict__ D, int N, int KD,
                        int a_mode, int b_mode, int use_tma, const __grid_constant__ CUtensorMap tmapA) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar, tbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *a_tile = smem;                    // K-major: KD/64 blocks of [128 rows][128 B]; MN-major: 2 blocks of [KD rows][128 B]
    uint8_t *b_tile = smem + 64 * 1024;        // K-major: KD/64 blocks of [N rows][128 B];   MN-major: N/64 blocks of [KD rows][128 B]
    const uint32_t a_blk = a_mode == 0 ? 128 * 128 : KD * 128;
    const uint32_t b_blk = b_mode == 0 ? N * 128 : KD * 128;
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&tbar, 1); mbar_fence_init(); }
    __syncthreads();
    if (use_tma && a_mode == 0) {
        if (tid == 0) {
            mbar_expect_tx(&tbar, 128 * KD * 2);
            for (int kb = 0; kb < KD / 64; ++kb) tma_load_2d(smem_u32(a_tile + kb * a_blk), &tmapA, kb * 64, 0, &tbar);
        }
        mbar_wait(&tbar, 0);
    } else {
        for (int i = tid; i < 128 * KD; i += 128) {
            const int m = i / KD, k = i % KD;
            const bf16 v = __float2bfloat16_rn(A[i]);
            if (a_mode == 0) *reinterpret_cast<bf16 *>(a_tile + (k / 64) * a_blk + sw128_off(m, (k % 64) / 8) + (k % 8) * 2) = v;
            else *reinterpret_cast<bf16 *>(a_tile + (m / 64) * a_blk + sw128_off(k, (m % 64) / 8) + (m % 8) * 2) = v;
        }
    }
    for (int i = tid; i < N * KD; i += 128) {
        const int n = i / KD, k = i % KD;
        const bf16 v = __float2bfloat16_rn(Bm[i]);
        if (b_mode == 0) *reinterpret_cast<bf16 *>(b_tile + (k / 64) * b_blk + sw128_off(n, (k % 64) / 8) + (k % 8) * 2) = v;
        else *reinterpret_cast<bf16 *>(b_tile + (n / 64) * b_blk + sw128_off(k, (n % 64) / 8) + (n % 8) * 2) = v;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = idesc_bf16(128, N, a_mode == 1, b_mode == 1);
        for (int ks = 0; ks < KD / 16; ++ks) {
            uint64_t ad, bd;
            if (a_mode == 0) ad = smem_desc_sw128(smem_u32(a_tile) + (ks / 4) * a_blk + (ks % 4) * 32, 0, 1024);
            else ad = smem_desc_sw128(smem_u32(a_tile) + ks * 2048, a_blk, 1024);
            if (b_mode == 0) bd = smem_desc_sw128(smem_u32(b_tile) + (ks / 4) * b_blk + (ks % 4) * 32, 0, 1024);
            else bd = smem_desc_sw128(smem_u32(b_tile) + ks * 2048, b_blk, 1024);
            mma_ss(tbase, ad, bd, idesc, ks > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tbase);
}

}  // namespace

// A_bf16 [128, KD] row-major (only read when use_tma != 0, through a 2-D tensor map with 128-byte swizzle)
extern "C" int lina_debug_umma_probe_sw128(const float *A, const float *B, float *D, const void *A_bf16, int N, int KD,
                                           int a_mode, int b_mode, int use_tma, void *stream) {
    LINA_REQUIRE(A && B && D, LINA_ERR_BAD_ARG, "umma_probe_sw128: null pointer");
    LINA_REQUIRE(N % 64 == 0 && N >= 64 && N <= 256 && KD % 64 == 0 && KD >= 64 && KD <= 128, LINA_ERR_BAD_ARG,
                 "umma_probe_sw128: need N in [64,256] %% 64, KD in {64,128}");
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    if (use_tma) {
        LINA_REQUIRE(A_bf16 != nullptr && a_mode == 0, LINA_ERR_BAD_ARG, "umma_probe_sw128: TMA needs A_bf16 and a_mode 0");
        const uint64_t dims[2] = {(uint64_t)KD, 128}, strides[2] = {2, (uint64_t)KD * 2};
        const uint32_t box[2] = {64, 128};
        int rc = lina_make_tmap_bf16(&tm, A_bf16, 2, dims, strides, box);
        if (rc) return rc;
    }
    const int smem = 140 * 1024;
    LINA_CUDA_OK(cudaFuncSetAttribute(umma_probe_sw128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_probe_sw128_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, KD, a_mode, b_mode, use_tma, tm);
    LINA_LAUNCH_OK("umma_probe_sw128_kernel");
    return LINA_OK;
}

// ---- third probe: cycles per tcgen05.mma for the shapes the GLA kernel issues (SW128 operands; data = garbage) ----
namespace {

template <bool ELECT>
__global__ void __launch_bounds__(128)
umma_timing_kernel(long long *__restrict__ out, int N, int a_tmem, int a_mn, int b_mn, int nmma, int same_d) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    // elect != 0: the issuing thread is chosen with elect.sync (ptxas then emits bare UTCHMMAs); elect == 0: `tid == 0`, under
    // which every MMA is wrapped in an elect / vote loop -- the ~97-cycle "issue cost" of round 1 was this loop
    bool issuer = tid == 0;
    if (ELECT) issuer = warp == 0 ? elect_one_sync() : false;      // compile-time choice: the region below is provably single-threaded
    if (issuer) {
        const uint32_t idesc = idesc_bf16(128, N, a_mn, b_mn);
        const uint32_t a_tile = smem_u32(smem), b_tile = smem_u32(smem + 64 * 1024);
        const uint64_t ad0 = a_mn ? smem_desc_sw128(a_tile, 8192, 1024) : smem_desc_sw128(a_tile, 0, 1024);
        const uint64_t bd0 = b_mn ? smem_desc_sw128(b_tile, 16384, 1024) : smem_desc_sw128(b_tile, 0, 1024);
        const uint64_t a_step = a_mn ? (2048 >> 4) : (32 >> 4), b_step = b_mn ? (2048 >> 4) : (32 >> 4);
        const uint32_t d1 = same_d ? 0u : 256u;
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            for (int i = 0; i < nmma; i += 4) {          // descriptors advanced by constants: 4 k-steps per round
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t d = tbase + ((j & 1) ? d1 : 0u);
                    if (a_tmem) mma_ts(d, tbase + 384 + j * 8, bd0 + j * b_step, idesc, 1);
                    else mma_ss(d, ad0 + j * a_step, bd0 + j * b_step, idesc, 1);
                }
            }
            const long long t1 = clock64();
            mma_commit(&bar);
            mbar_wait(&bar, rep & 1);
            const long long t2 = clock64();
            out[rep * 2 + 0] = t1 - t0;      // issue time
            out[rep * 2 + 1] = t2 - t0;      // issue + completion
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tbase);
}

}  // namespace

// out[6]: (issue cycles, total cycles) x 3 repetitions of `nmma` back-to-back M=128 x N x K=16 bf16 MMAs
extern "C" int lina_debug_umma_timing(long long *out, int N, int a_tmem, int a_mn, int b_mn, int nmma, int same_d,
                                      void *stream) {
    LINA_REQUIRE(out && N % 16 == 0 && N >= 16 && N <= 256 && nmma > 0 && nmma <= 4096, LINA_ERR_BAD_ARG, "umma_timing: bad argument");
    const int smem = 160 * 1024;
    LINA_CUDA_OK(cudaFuncSetAttribute(umma_timing_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    LINA_CUDA_OK(cudaFuncSetAttribute(umma_timing_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if ((same_d >> 1) & 1) umma_timing_kernel<true><<<1, 128, smem, (cudaStream_t)stream>>>(out, N, a_tmem, a_mn, b_mn, nmma, same_d & 1);
    else umma_timing_kernel<false><<<1, 128, smem, (cudaStream_t)stream>>>(out, N, a_tmem, a_mn, b_mn, nmma, same_d & 1);
    LINA_LAUNCH_OK("umma_timing_kernel");
    return LINA_OK;
}
