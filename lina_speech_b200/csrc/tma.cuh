// Host side of TMA: CUtensorMap construction through the driver entry point (no link-time libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"

typedef CUresult (*lina_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                         const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                         CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                         CUtensorMapFloatOOBfill);

static inline lina_encode_tiled_fn lina_get_encode_tiled() {
    static lina_encode_tiled_fn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = (lina_encode_tiled_fn)p;
    }
    return fn;
}

// bf16 tensor of `rank` dims (innermost first), 128-byte swizzle, zero fill out of bounds.
// dims[i] elements, strides_bytes[i] for i >= 1 (stride of dim i), box[i] elements.
static inline int lina_make_tmap_bf16(CUtensorMap *m, const void *base, int rank, const uint64_t *dims,
                                      const uint64_t *strides_bytes, const uint32_t *box) {
    lina_encode_tiled_fn enc = lina_get_encode_tiled();
    LINA_REQUIRE(enc != nullptr, LINA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gd[5], gs[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i + 1];
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), gd, gs, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LINA_REQUIRE(r == CUDA_SUCCESS, LINA_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return LINA_OK;
}
