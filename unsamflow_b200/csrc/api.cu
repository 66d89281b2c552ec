// api.cu -- the extern "C" boundary declared in include/unsamflow_b200.h.
// Validates arguments, dispatches on dtype, never allocates, never synchronises.
#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace usf {

int corr_fwd_f32(const float *, const float *, float *, int, int, int, int, unsigned, float, cudaStream_t);
int corr_bwd_f32(const float *, const float *, const float *, float *, float *, int, int, int, int,
                 cudaStream_t);
int warp_fwd_f32(const float *, const float *, float *, int, int, int, int, int, unsigned, cudaStream_t);
int warp_bwd_f32(const float *, const float *, const float *, float *, float *, int, int, int, int, int,
                 unsigned, void *, size_t, cudaStream_t);
size_t warp_bwd_workspace_bytes(int B, int H, int W);
size_t warp_bwd_workspace_clean_bytes(int B, int H, int W);
int corr_fwd_bf16(const void *, const void *, void *, int, int, int, int, unsigned, float, cudaStream_t);
int corr_fwd_bf16_tiled(const void *, const void *, void *, int, int, int, int, unsigned, float, cudaStream_t);
int corr_bwd_bf16(const void *, const void *, const void *, void *, void *, int, int, int, int, cudaStream_t);
int corr_bwd_bf16_tiled(const void *, const void *, const void *, void *, void *, int, int, int, int, cudaStream_t);
int warp_fwd_bf16(const void *, const float *, void *, int, int, int, int, int, unsigned, cudaStream_t);
int warp_bwd_bf16(const void *, const float *, const void *, void *, float *, int, int, int, int, int, unsigned,
                  void *, size_t, cudaStream_t);

int corr_fwd_small_f32(const float *, const float *, float *, int, int, int, int, unsigned, float, cudaStream_t);
int corr_bwd_small_f32(const float *, const float *, const float *, float *, float *, int, int, int, int,
                       cudaStream_t);
int corr_fwd_small_bf16(const void *, const void *, void *, int, int, int, int, unsigned, float, cudaStream_t);
int corr_bwd_small_bf16(const void *, const void *, const void *, void *, void *, int, int, int, int, cudaStream_t);
int occ_bidirection(const float *, const float *, float *, int, int, int, float, float, unsigned, cudaStream_t);
size_t occ_backward_workspace_bytes(int B, int H, int W);
int occ_backward(const float *, float *, float *, int, int, int, float, void *, cudaStream_t);
int segpool_max_segments();
int segpool_fwd(const void *, const int *, void *, int *, int, int, int, int, int, cudaStream_t);
int segpool_bwd(const void *, const int *, const int *, void *, int, int, int, int, int, cudaStream_t);
bool corr_small_applies(int H, int W);  // whole image <= 256 pixels: corr_small.cu

long kDirectMaxPixels = [] {
    const char *e = getenv("USF_DIRECT_MAX_PIXELS");
    return e ? atol(e) : 4096L;
}();

static thread_local size_t tl_out_bstride = 0;  // 0: dense [B,81,H,W]
size_t corr_out_bstride(int H, int W) { return tl_out_bstride ? tl_out_bstride : (size_t)kK * H * W; }

int sm_count() {
    static std::mutex mu;
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    std::lock_guard<std::mutex> lock(mu);
    if (cache[dev] == 0) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
        cache[dev] = sms;
    }
    return cache[dev];
}

bool pdl_enabled() {
    static const bool on = [] {
        const char *e = getenv("USF_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

static thread_local int tl_cuda_error = 0;
static thread_local unsigned long long tl_launches = 0;

void note_cuda_error(cudaError_t e) { tl_cuda_error = (int)e; }

cudaError_t ensure_smem(const void *kernel, int smem_bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, int> done;  // (kernel, device) -> bytes already granted
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(kernel, dev);
    auto it = done.find(key);
    if (it != done.end() && it->second >= smem_bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    done[key] = smem_bytes;
    return cudaSuccess;
}

int note_launch(cudaError_t e) {
    if (e != cudaSuccess) {
        tl_cuda_error = (int)e;
        return USF_ERR_CUDA;
    }
    ++tl_launches;
    return USF_OK;
}

static int check_dims(int B, int C, int H, int W, int chan_max) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return USF_ERR_BAD_SHAPE;
    const long long plane = (long long)H * W;
    const int cm = C > chan_max ? C : chan_max;
    // 32-bit index bound on the largest tensor the call touches
    if ((long long)B * cm * plane > 2147483647LL) return USF_ERR_BAD_SHAPE;
    return USF_OK;
}

}  // namespace usf

using namespace usf;

extern "C" {

int usf_abi_version(void) { return USF_ABI_VERSION; }

const char *usf_status_string(int status) {
    switch (status) {
        case USF_OK: return "ok";
        case USF_ERR_BAD_SHAPE: return "bad shape (non-positive dim, H or W < 2 for warp, or > 2^31-1 elements)";
        case USF_ERR_BAD_ARG: return "bad argument (null pointer, unknown dtype/pad/flag, or max_disp != 4)";
        case USF_ERR_WORKSPACE: return "workspace missing, misaligned or too small (see usf_*_workspace_bytes)";
        case USF_ERR_CUDA: return "CUDA call failed (see usf_last_cuda_error)";
        case USF_ERR_UNSUPPORTED: return "not implemented in this build";
        default: return "unknown status";
    }
}

int usf_last_cuda_error(void) { return tl_cuda_error; }
unsigned long long usf_launch_count(void) { return tl_launches; }
void usf_reset_launch_count(void) { tl_launches = 0; }

int usf_corr_fwd(const void *x1, const void *x2, void *out, int B, int C, int H, int W, int max_disp,
                 int dtype, unsigned flags, float leaky_slope, void *stream) {
    if (!x1 || !x2 || !out || max_disp != kMaxDisp) return USF_ERR_BAD_ARG;
    if (flags & ~USF_FLAG_LEAKY_RELU) return USF_ERR_BAD_ARG;
    if (int st = check_dims(B, C, H, W, kK)) return st;
    if (corr_small_applies(H, W)) {
        if (dtype == USF_F32)
            return corr_fwd_small_f32((const float *)x1, (const float *)x2, (float *)out, B, C, H, W, flags,
                                      leaky_slope, (cudaStream_t)stream);
        if (dtype == USF_BF16)
            return corr_fwd_small_bf16(x1, x2, out, B, C, H, W, flags, leaky_slope, (cudaStream_t)stream);
        return USF_ERR_BAD_ARG;
    }
    if (dtype == USF_F32)
        return corr_fwd_f32((const float *)x1, (const float *)x2, (float *)out, B, C, H, W, flags, leaky_slope,
                            (cudaStream_t)stream);
    if (dtype == USF_BF16) {
        if ((long)B * H * W > kDirectMaxPixels) {  // tiled TMA kernel when the shape allows (W % 8 == 0)
            const int st = corr_fwd_bf16_tiled(x1, x2, out, B, C, H, W, flags, leaky_slope, (cudaStream_t)stream);
            if (st != USF_ERR_UNSUPPORTED) return st;
        }
        return corr_fwd_bf16(x1, x2, out, B, C, H, W, flags, leaky_slope, (cudaStream_t)stream);
    }
    return USF_ERR_BAD_ARG;
}

int usf_corr_fwd_into(const void *x1, const void *x2, void *out, long long out_batch_stride, int B, int C, int H,
                      int W, int max_disp, int dtype, unsigned flags, float leaky_slope, void *stream) {
    if (out_batch_stride < (long long)kK * H * W) return USF_ERR_BAD_SHAPE;
    if ((long long)B * out_batch_stride > 2147483647LL) return USF_ERR_BAD_SHAPE;
    tl_out_bstride = (size_t)out_batch_stride;
    const int st = usf_corr_fwd(x1, x2, out, B, C, H, W, max_disp, dtype, flags, leaky_slope, stream);
    tl_out_bstride = 0;
    return st;
}

int usf_corr_bwd(const void *x1, const void *x2, const void *gout, void *gx1, void *gx2, int B, int C,
                 int H, int W, int max_disp, int dtype, unsigned flags, void *stream) {
    if (!x1 || !x2 || !gout || max_disp != kMaxDisp || flags != 0) return USF_ERR_BAD_ARG;
    if (int st = check_dims(B, C, H, W, kK)) return st;
    if (!gx1 && !gx2) return USF_OK;
    if (corr_small_applies(H, W)) {
        if (dtype == USF_F32)
            return corr_bwd_small_f32((const float *)x1, (const float *)x2, (const float *)gout, (float *)gx1,
                                      (float *)gx2, B, C, H, W, (cudaStream_t)stream);
        if (dtype == USF_BF16) return corr_bwd_small_bf16(x1, x2, gout, gx1, gx2, B, C, H, W, (cudaStream_t)stream);
        return USF_ERR_BAD_ARG;
    }
    if (dtype == USF_F32)
        return corr_bwd_f32((const float *)x1, (const float *)x2, (const float *)gout, (float *)gx1,
                            (float *)gx2, B, C, H, W, (cudaStream_t)stream);
    if (dtype == USF_BF16) {
        if ((long)B * H * W > kDirectMaxPixels) {  // tiled TMA kernel when the shape allows (W % 8 == 0)
            const int st = corr_bwd_bf16_tiled(x1, x2, gout, gx1, gx2, B, C, H, W, (cudaStream_t)stream);
            if (st != USF_ERR_UNSUPPORTED) return st;
        }
        return corr_bwd_bf16(x1, x2, gout, gx1, gx2, B, C, H, W, (cudaStream_t)stream);
    }
    return USF_ERR_BAD_ARG;
}

int usf_warp_fwd(const void *x, const float *flow, void *out, int B, int C, int H, int W, int pad_mode,
                 int dtype, unsigned flags, void *stream) {
    if (!x || !flow || !out) return USF_ERR_BAD_ARG;
    if (pad_mode != USF_PAD_BORDER && pad_mode != USF_PAD_ZEROS) return USF_ERR_BAD_ARG;
    if (flags & ~USF_FLAG_COORD_RECIPROCAL) return USF_ERR_BAD_ARG;
    if (int st = check_dims(B, C, H, W, 2)) return st;
    if (H < 2 || W < 2) return USF_ERR_BAD_SHAPE;  // the reference divides by (size-1): utils/warp_utils.py:21-22
    if (dtype == USF_F32)
        return warp_fwd_f32((const float *)x, flow, (float *)out, B, C, H, W, pad_mode, flags,
                            (cudaStream_t)stream);
    if (dtype == USF_BF16) return warp_fwd_bf16(x, flow, out, B, C, H, W, pad_mode, flags, (cudaStream_t)stream);
    return USF_ERR_BAD_ARG;
}

size_t usf_warp_bwd_workspace_bytes(int B, int C, int H, int W, unsigned flags) {
    (void)C;
    if (B <= 0 || H <= 0 || W <= 0 || (flags & USF_FLAG_WARP_BWD_ATOMIC)) return 0;
    return warp_bwd_workspace_bytes(B, H, W);
}

size_t usf_warp_bwd_workspace_clean_bytes(int B, int C, int H, int W, unsigned flags) {
    (void)C;
    if (B <= 0 || H <= 0 || W <= 0 || (flags & USF_FLAG_WARP_BWD_ATOMIC)) return 0;
    return warp_bwd_workspace_clean_bytes(B, H, W);
}

int usf_warp_bwd(const void *x, const float *flow, const void *gout, void *gx, float *gflow, int B, int C,
                 int H, int W, int pad_mode, int dtype, unsigned flags, void *workspace, size_t workspace_bytes,
                 void *stream) {
    if (!x || !flow || !gout) return USF_ERR_BAD_ARG;
    if (pad_mode != USF_PAD_BORDER && pad_mode != USF_PAD_ZEROS) return USF_ERR_BAD_ARG;
    if (flags & ~(USF_FLAG_COORD_RECIPROCAL | USF_FLAG_WARP_BWD_ATOMIC | USF_FLAG_WORKSPACE_CLEAN)) return USF_ERR_BAD_ARG;
    if (int st = check_dims(B, C, H, W, 2)) return st;
    if (H < 2 || W < 2) return USF_ERR_BAD_SHAPE;
    if (!gx && !gflow) return USF_OK;
    if (gx && (long long)B * H * W > (1LL << 29)) return USF_ERR_BAD_SHAPE;  // 4*N tap bound stays in int
    if (dtype == USF_F32)
        return warp_bwd_f32((const float *)x, flow, (const float *)gout, (float *)gx, gflow, B, C, H, W,
                            pad_mode, flags, workspace, workspace_bytes, (cudaStream_t)stream);
    if (dtype == USF_BF16)
        return warp_bwd_bf16(x, flow, gout, gx, gflow, B, C, H, W, pad_mode, flags, workspace, workspace_bytes,
                             (cudaStream_t)stream);
    return USF_ERR_BAD_ARG;
}

int usf_occ_bidirection(const float *flow12, const float *flow21, float *occ, int B, int H, int W, float scale,
                        float bias, unsigned flags, void *stream) {
    if (!flow12 || !flow21 || !occ) return USF_ERR_BAD_ARG;
    if (flags & ~USF_FLAG_COORD_RECIPROCAL) return USF_ERR_BAD_ARG;
    if (int st = check_dims(B, 2, H, W, 2)) return st;
    if (H < 2 || W < 2) return USF_ERR_BAD_SHAPE;  // flow_warp divides by (size-1): utils/warp_utils.py:21-22
    return occ_bidirection(flow12, flow21, occ, B, H, W, scale, bias, flags, (cudaStream_t)stream);
}

size_t usf_occ_backward_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return occ_backward_workspace_bytes(B, H, W);
}

int usf_occ_backward(const float *flow21, float *occ, float *range_map, int B, int H, int W, float th,
                     void *workspace, size_t workspace_bytes, void *stream) {
    if (!flow21 || !occ) return USF_ERR_BAD_ARG;
    if (int st = check_dims(B, 2, H, W, 2)) return st;
    if (H > (1 << 24) || W > (1 << 24)) return USF_ERR_BAD_SHAPE;  // pixel coordinates must be exact in fp32
    if (!workspace || workspace_bytes < occ_backward_workspace_bytes(B, H, W) || ((uintptr_t)workspace & 7))
        return USF_ERR_WORKSPACE;
    return occ_backward(flow21, occ, range_map, B, H, W, th, workspace, (cudaStream_t)stream);
}

int usf_segpool_max_segments(void) { return segpool_max_segments(); }

int usf_segpool_fwd(const void *feat, const int *seg, void *out, int *argmax, int B, int C, int H, int W,
                    int num_segments, int dtype, void *stream) {
    if (!feat || !seg || !out) return USF_ERR_BAD_ARG;
    if (dtype != USF_F32 && dtype != USF_BF16) return USF_ERR_BAD_ARG;
    if (int st = check_dims(B, C, H, W, 1)) return st;
    if (num_segments < 1 || num_segments > segpool_max_segments()) return USF_ERR_BAD_SHAPE;
    if ((long long)B * C * num_segments > 2147483647LL) return USF_ERR_BAD_SHAPE;
    return segpool_fwd(feat, seg, out, argmax, B, C, H * W, num_segments, dtype, (cudaStream_t)stream);
}

int usf_segpool_bwd(const void *gout, const int *seg, const int *argmax, void *gfeat, int B, int C, int H, int W,
                    int num_segments, int dtype, void *stream) {
    if (!gout || !seg || !argmax || !gfeat) return USF_ERR_BAD_ARG;
    if (dtype != USF_F32 && dtype != USF_BF16) return USF_ERR_BAD_ARG;
    if (int st = check_dims(B, C, H, W, 1)) return st;
    if (num_segments < 1 || num_segments > segpool_max_segments()) return USF_ERR_BAD_SHAPE;
    if ((long long)B * C * num_segments > 2147483647LL) return USF_ERR_BAD_SHAPE;
    return segpool_bwd(gout, seg, argmax, gfeat, B, C, H * W, num_segments, dtype, (cudaStream_t)stream);
}

}
