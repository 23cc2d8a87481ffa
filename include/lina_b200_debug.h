/* liblina_b200_debug.so -- bring-up probes that pin the tcgen05 descriptor conventions on hardware (tests/test_umma_probe_gpu.py,
 * profiles/probe_m64.py, profiles/umma_timing.py).  NOT part of the product library: built from csrc/debug/ (plus a traced
 * copy of the GLA chunk kernel and the product objects it calls into) into its own shared object, declared here and not in
 * lina_b200.h. */
#ifndef LINA_B200_DEBUG_H
#define LINA_B200_DEBUG_H
#ifdef __cplusplus
extern "C" {
#endif
/* ---------------------------------------------------------------------------------------------
 * Debug / bring-up: one-CTA tcgen05 GEMM D[128,N] = A[128,KD] * B[N,KD]^T (fp32 in, bf16 math) with the
 * operand placements of the GLA kernel (a_mode: 0 smem K-major, 1 smem MN-major, 2 TMEM; b_mode: 0 / 1).
 * `swap` exchanges the descriptor's leading/stride byte offsets.  Not part of the reference's API.
 * ------------------------------------------------------------------------------------------- */
int lina_debug_umma_probe(const float *A, const float *B, float *D, int N, int KD, int a_mode, int b_mode,
                          int swap, void *stream);
/* Round-2 bring-up (not yet run on hardware): M x N x 16 MMAs with M in {64,128}, no-swizzle K-major operands; D receives the
 * RAW accumulator tile [128 TMEM lanes][N] (cells the MMA did not write hold -12345) so the M = 64 lane mapping can be read. */
int lina_debug_umma_probe_m(const float *A, const float *B, float *D, int M, int N, int KD, void *stream);
/* Same with 128-byte-swizzled operands (a_mode / b_mode: 0 K-major, 1 MN-major); use_tma != 0 loads A from
 * A_bf16 [128,KD] through a 2-D tensor map with CU_TENSOR_MAP_SWIZZLE_128B instead of writing it by hand. */
int lina_debug_umma_probe_sw128(const float *A, const float *B, float *D, const void *A_bf16, int N, int KD,
                                int a_mode, int b_mode, int use_tma, void *stream);
/* Cycles of `nmma` back-to-back M=128 x N x 16 bf16 MMAs issued by one thread (A from TMEM / smem K-major /
 * smem MN-major, B K-/MN-major, same or alternating accumulator): out[6] = (issue, issue+completion) x 3 reps. */
int lina_debug_umma_timing(long long *out, int N, int a_tmem, int a_mn, int b_mn, int nmma, int same_d, void *stream);
/* The tcgen05 GLA kernels (K = 256, bf16) compiled WITH their clock64 timeline (-DLINA_GLA_TRACE; the product library's copy has
 * none): trace[6 roles][64 items][4 events] (int64) of the first CTA -- profiles/trace_gla_chunk.py, trace_gla_pregated.py and
 * trace_gla_pair.py print it.  The pre-gated entry runs the CTA-pair kernel unless lina_debug_set_variant(9, 1) was called ON
 * THIS LIBRARY. */
int lina_debug_set_variant(int key, int value);
int lina_debug_gla_chunk_trace(const void *q, const void *k, const void *v, const void *gk, void *o, int B,
                               int H, int T, int K, int V, float scale, long long *trace, void *stream);
int lina_debug_gla_pregated_trace(const void *qg, const void *kg, const void *v, const float *decay, void *o, int B,
                                  int H, int T, int K, int V, long long *trace, void *stream);
#ifdef __cplusplus
}
#endif
#endif /* LINA_B200_DEBUG_H */
