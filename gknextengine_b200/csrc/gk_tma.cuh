// gk_tma.cuh — minimal TMA (cp.async.bulk.tensor) helpers for the stencil kernels: 2-D tiled tensor maps over the storage
// images, one elected thread issues the bulk copies, everybody waits on an mbarrier. Out-of-bounds box elements (negative or
// >= extent coordinates) are zero-filled by the hardware, which is exactly Vulkan's robust image access the reference relies on
// for its halo reads (Process.ReProject.comp.slang:137-150, Process.DenoiseJBF.comp.slang:103-118).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gktma {

// Host: encode a 2-D map. `elem` is CU_TENSOR_MAP_DATA_TYPE_UINT64 for fp16x4 / RG32F pixels, UINT32 for object ids.
// Returns false when the driver entry point is missing or the image violates TMA's 16-byte stride rule (odd widths).
inline bool encode_2d(CUtensorMap* out, void* base, CUtensorMapDataType elem, uint32_t elem_bytes, uint64_t width, uint64_t height, uint32_t box_w,
                      uint32_t box_h) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
        fn = reinterpret_cast<EncodeFn>(p);
    }
    // 16-byte rules of cuTensorMapEncodeTiled; in addition the kernel must start every box on a 16-byte boundary (see gk_post.cu)
    if ((width * elem_bytes) % 16 != 0 || (box_w * elem_bytes) % 16 != 0 || box_w > 256 || box_h > 256) return false;
    cuuint64_t dims[2] = {width, height};
    cuuint64_t strides[1] = {width * elem_bytes};
    cuuint32_t box[2] = {box_w, box_h};
    cuuint32_t estr[2] = {1, 1};
    return fn(out, elem, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
                 "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
#endif

}  // namespace gktma
