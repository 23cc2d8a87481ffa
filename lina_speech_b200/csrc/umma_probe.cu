// Single-CTA tcgen05 probe: D[128,N] = A[128,KD] * B[N,KD]^T in bf16 -> fp32, with every operand
// placement the GLA kernel uses (smem K-major, smem MN-major, A from TMEM).  Exists to pin the
// descriptor conventions of sm100.cuh on real hardware (tests/test_umma_probe_gpu.py).
#include "common.cuh"
#include "sm100.cuh"

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
