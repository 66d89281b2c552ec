// tma.cuh -- TMA (cp.async.bulk.tensor) + mbarrier helpers and the host-side tensor-map encoder.
//
// Feature maps are NCHW fp32; a 4-D tensor map (W, H, C, B) lets one thread fetch a
// [channels][rows][cols] box with its halo in one instruction.  Boxes may start at negative
// coordinates or run past the image / channel count: the hardware zero-fills, which is exactly
// F.pad(zeros) of models/correlation_native.py:16 and the ragged channel tail.
// Requirements (else the callers fall back to the cp.async kernels): W % 4 == 0 (global strides
// are multiples of 16 bytes) and a 16-byte aligned base pointer.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace usf {

// ---- host ------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled is fetched through the runtime so the library does not link libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// NCHW fp32 tensor [B,C,H,W] -> map with box [1][bc][bh][bw].  Returns false if TMA cannot be used.
// USF_DISABLE_TMA=1 in the environment forces the cp.async kernels (A/B measurements, debugging).
inline bool tma_enabled() {
    static const bool on = [] {
        const char *e = getenv("USF_DISABLE_TMA");
        return !(e && e[0] == '1');
    }();
    return on;
}

inline bool make_nchw_map(CUtensorMap *map, const float *base, int B, int C, int H, int W, int bw, int bh,
                          int bc) {
    if (!tma_enabled()) return false;
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || (W % 4) != 0 || ((uintptr_t)base & 15) != 0) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
    const cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// bf16 NCHW tensor, box [bc][bh][bw]: global strides must be multiples of 16 bytes -> W % 8 == 0
inline bool make_nchw_map_bf16(CUtensorMap *map, const void *base, int B, int C, int H, int W, int bw, int bh,
                               int bc) {
    if (!tma_enabled()) return false;
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || (W % 8) != 0 || ((uintptr_t)base & 15) != 0) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 2, (cuuint64_t)H * W * 2, (cuuint64_t)C * H * W * 2};
    const cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- device ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(s32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "USF_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra USF_DONE;\n\t"
        "bra USF_WAIT;\n"
        "USF_DONE:\n\t}\n" ::"r"(s32(bar)),
        "r"(parity)
        : "memory");
}
// box load: coordinates are (x, y, c, b), innermost first
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y,
                                            int c, int b) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::
            "r"(s32(smem_dst)),
        "l"((uint64_t)map), "r"(s32(bar)), "r"(x), "r"(y), "r"(c), "r"(b)
        : "memory");
}
// 2-D box load: coordinates are (column, row)
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
            "r"(s32(smem_dst)),
        "l"((uint64_t)map), "r"(s32(bar)), "r"(x), "r"(y)
        : "memory");
}
// 1-D bulk copy global -> shared (no tensor map): 16-byte aligned source, destination and size
__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
                     "r"(s32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(s32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"((uint64_t)map) : "memory");
}

}  // namespace usf
