// common.cuh -- shared device helpers for the corr / warp kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

#include "../../include/unsamflow_b200.h"

namespace usf {

constexpr int kMaxDisp = 4;            // PWC-Lite search range (models/pwclite.py:182)
constexpr int kD = 2 * kMaxDisp + 1;   // 9
constexpr int kK = kD * kD;            // 81 displacement channels
// B*H*W at or below which the direct (no shared memory) corr kernels are used; override with
// USF_DIRECT_MAX_PIXELS for A/B runs
extern long kDirectMaxPixels;

// ---- host-side bookkeeping (api.cu) ---------------------------------------
int note_launch(cudaError_t e);        // counts a launch, records e; returns usf_status
void note_cuda_error(cudaError_t e);
// Opt a kernel into `smem_bytes` of dynamic shared memory (+ max shared carve-out) once per (kernel, device).
// cudaFuncSetAttribute is per device, so a process that drives several GPUs is handled too.
cudaError_t ensure_smem(const void *kernel, int smem_bytes);

// Output batch stride of the correlation forward, in elements: 81*H*W for a dense [B,81,H,W] cost volume, larger when
// the caller points `out` at a channel slice of a wider tensor (usf_corr_fwd_into: the estimator's concat buffer,
// models/pwclite.py:364-366).  api.cu parks the requested value in a thread-local for the duration of the call.
size_t corr_out_bstride(int H, int W);

// ---- launches ---------------------------------------------------------------
// Every hot-path kernel is launched with programmatic stream serialization and starts with griddepcontrol.wait:
// the semantics are those of plain stream order (the wait returns when the preceding kernel of the stream has
// completed and flushed), but the launch itself -- grid setup, parameter and tensor-map upload, CTA dispatch -- no
// longer waits for the predecessor to drain, which is a microsecond or two per dependent kernel in a chain of
// ~30.  USF_PDL=0 launches without the attribute (the wait is then a no-op).
bool pdl_enabled();  // api.cu
// SMs of the current device (cudaDevAttrMultiProcessorCount, cached per device; 148 on B200): grid sizing
int sm_count();      // api.cu
// Tried in round 2 and dropped: griddepcontrol.launch_dependents right after the wait (the next kernel's CTAs are
// then dispatched into freed SM slots and sit in their own wait while this grid drains).  Safe by construction --
// every kernel's first statement is the wait -- but the step got 3 % slower (0.598 -> 0.614 ms): the parked CTAs hold
// registers and shared memory that the two concurrently running direction streams would otherwise use.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t usf_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- cp.async (LDGSTS) ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// 16-byte copy, bypassing L1; src_bytes == 0 zero-fills the destination.
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem),
                 "r"(src_bytes)
                 : "memory");
}
// 4-byte copy; src_bytes == 0 zero-fills.
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem),
                 "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// storage-type agnostic scalar access: feature tensors are fp32 or bf16, arithmetic is always fp32
typedef __nv_bfloat16 bf16;
__device__ __forceinline__ float ldf(const float *p) { return __ldg(p); }
__device__ __forceinline__ float ldf(const bf16 *p) { return __bfloat162float(__ldg(p)); }
__device__ __forceinline__ void stf(float *p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16 *p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void atomic_addf(float *p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_addf(bf16 *p, float v) { atomicAdd(p, __float2bfloat16_rn(v)); }

__device__ __forceinline__ float4 lds128(const float *p) { return *reinterpret_cast<const float4 *>(p); }

// Release of a shared-memory stage back to the TMA producer.  The fence makes every lane's outstanding
// shared-memory reads of the stage COMPLETE before lane 0 arrives on the "empty" barrier: an arrive that merely
// follows the LDS instructions in program order does not wait for their data, and ptxas is free to sink the
// FFMAs that consume them below the arrive (it does, in the 3-CTA/SM backward) -- the producer's next TMA
// write then races with loads still in flight (seen as a few non-reproducible runs out of hundreds).
__device__ __forceinline__ void release_stage_reads() {
    __threadfence_block();  // MEMBAR.CTA: waits for the outstanding loads
    __syncwarp();
}

__device__ __forceinline__ void st_global_cs(float *p, float v) {
    asm volatile("st.global.cs.f32 [%0], %1;\n" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_global_cs4(float *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

}  // namespace usf
