// Host-side helper shared by the kernels that use TMA: cuTensorMapEncodeTiled through the runtime's
// driver entry point (no -lcuda needed).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace tma_host {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
        if (q != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// dims / box: innermost first; strides_bytes: rank-1 entries (dimension 0 is dense)
inline bool encode(CUtensorMap* m, CUtensorMapDataType dt, const void* base, int rank, const cuuint64_t* dims,
                   const cuuint64_t* strides_bytes, const cuuint32_t* box, CUtensorMapSwizzle sw) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return false;
    // The driver call needs a current context.  A host thread that has not issued a runtime call yet (autograd's
    // worker thread running a backward as its first CUDA work) has none: bind the primary context then — and only
    // then, because cudaFree is not allowed while a stream capture is in progress (a capturing thread has one).
    typedef CUresult (*CtxGetCurrentFn)(CUcontext*);
    static CtxGetCurrentFn ctx_get = []() -> CtxGetCurrentFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuCtxGetCurrent", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<CtxGetCurrentFn>(p);
    }();
    CUcontext cur = nullptr;
    if (!ctx_get || ctx_get(&cur) != CUDA_SUCCESS || cur == nullptr) cudaFree(nullptr);
    cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    return enc(m, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
               sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tma_host
