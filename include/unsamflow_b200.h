/*
 * unsamflow_b200.h -- C ABI of the B200-native correlation / flow_warp hot path.
 *
 * This is the drop-in boundary for UnSAMFlow's PWC-Lite hot path.  Each entry
 * point replaces one piece of the reference (paths relative to the reference
 * root; see INTEGRATION.md for the binding a maintainer adds on the reference
 * side):
 *
 *   usf_corr_fwd   correlation_cuda.forward  (models/correlation_package/
 *                  correlation_cuda.cc:10-86, pybind :167-170), i.e. what
 *                  CorrelationFunction.forward calls (correlation.py:24); same
 *                  result as models/correlation_native.py:13-23.
 *   usf_corr_bwd   correlation_cuda.backward (correlation_cuda.cc:88-165), called
 *                  from CorrelationFunction.backward (correlation.py:56).
 *   usf_warp_fwd   utils/warp_utils.py:97-106 flow_warp (mesh_grid :7-13 +
 *                  norm_grid :16-23 + F.grid_sample) forward.
 *   usf_warp_bwd   autograd backward of the same (ATen grid_sampler_2d_backward
 *                  chained through norm_grid).
 *   usf_occ_bidirection   utils/warp_utils.py:109-117 get_occu_mask_bidirection
 *                  (flow_warp(flow21, flow12, pad="zeros") + the forward-backward
 *                  consistency test), one kernel.
 *   usf_segpool_fwd / _bwd  models/pwclite.py:317-340: the segment max-pooling of the mask-feature branch
 *                  (one_hot x feature -> amax -> un-pool) of the "+mf" configs, and its gradient.
 *   usf_occ_backward      utils/warp_utils.py:26-94,120-126 get_occu_mask_backward
 *                  (get_corresponding_map's 4-tap forward splat + clamp + compare);
 *                  the mask losses/flow_loss.py:99-113 uses with occ_from_back.
 *
 * Conventions (differences from the reference ABI are deliberate):
 *   - plain pointers and sizes; no at::Tensor crosses the boundary;
 *   - all tensors are DEVICE pointers, NCHW, contiguous; `dtype` selects the
 *     storage type of feature tensors (flow and its gradient are always fp32);
 *   - the caller owns every buffer, including the workspace; nothing is allocated,
 *     freed, resized or zero-filled behind the caller's back, except that
 *     usf_warp_bwd fully overwrites gx / gflow (no pre-zeroing needed);
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work, never
 *     synchronise, and are CUDA-graph capturable and re-entrant;
 *   - return value: USF_OK (0) or a negative usf_status; CUDA launch failures are
 *     returned as USF_ERR_CUDA with the cudaError_t available from
 *     usf_last_cuda_error().  (The reference returns 0 on failure and prints,
 *     correlation_cuda_kernel.cu:380-387; the Python shim turns non-zero into
 *     RuntimeError like correlation_cuda.cc:80-82 does.)
 */
#ifndef UNSAMFLOW_B200_H_
#define UNSAMFLOW_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define USF_ABI_VERSION 3

typedef enum usf_status {
    USF_OK = 0,
    USF_ERR_BAD_SHAPE = -1,   /* non-positive dims, H or W == 1 for warp, index > 2^31-1 */
    USF_ERR_BAD_ARG = -2,     /* null pointer, unknown dtype / pad mode / flag, md != 4 */
    USF_ERR_WORKSPACE = -3,   /* workspace missing, misaligned or smaller than usf_*_workspace_bytes */
    USF_ERR_CUDA = -4,        /* a CUDA runtime call failed; see usf_last_cuda_error() */
    USF_ERR_UNSUPPORTED = -5  /* valid request this build does not implement (none at ABI version 3) */
} usf_status;

/* storage type of feature tensors; arithmetic is fp32 either way (bf16: one rounding per output element) */
typedef enum usf_dtype { USF_F32 = 0, USF_BF16 = 1 } usf_dtype;

/* padding_mode of flow_warp (utils/warp_utils.py:97: pad="border" | "zeros") */
typedef enum usf_pad { USF_PAD_BORDER = 0, USF_PAD_ZEROS = 1 } usf_pad;

/* flags, OR-able */
#define USF_FLAG_NONE 0u
/* corr fwd: apply LeakyReLU(slope) to the output in the epilogue
 * (models/pwclite.py:308 does it in place right after the op). */
#define USF_FLAG_LEAKY_RELU 1u
/* warp: compute the normalised coordinate as v*2 * (1/(size-1)) instead of
 * v*2 / (size-1).  ATen's CUDA `tensor / python_scalar` multiplies by the
 * reciprocal while its CPU kernel divides; the default follows the CPU reference
 * (the oracle), this flag follows the reference when it runs on a GPU. */
#define USF_FLAG_COORD_RECIPROCAL 2u
/* warp bwd: accumulate gx with global float atomics (red.global.add.f32, what ATen
 * does) instead of the default atomic-free tap-list gather.  For A/B measurements. */
#define USF_FLAG_WARP_BWD_ATOMIC 4u
/* warp bwd: the caller guarantees that the first usf_warp_bwd_workspace_clean_bytes() bytes of
 * `workspace` are zero -- after a one-time memset, or after any earlier usf_warp_bwd call on the
 * same workspace, which always leaves them zeroed -- so the per-call memset node is skipped.
 * For callers that keep one workspace per call site alive across steps (CUDA-graph replays). */
#define USF_FLAG_WORKSPACE_CLEAN 8u

int usf_abi_version(void);
const char *usf_status_string(int status);
/* cudaError_t of the most recent USF_ERR_CUDA on the calling thread (0 if none). */
int usf_last_cuda_error(void);
/* Number of kernels launched by this library on the calling thread since the
 * last usf_reset_launch_count(); used by bench.py for "gpu_launches". */
unsigned long long usf_launch_count(void);
void usf_reset_launch_count(void);

/*
 * Cost volume.  out[b, (dy+md)*(2md+1) + (dx+md), y, x] =
 *     (1/C) * sum_c x1[b,c,y,x] * x2[b,c,y+dy,x+dx],   x2 == 0 outside the image.
 * x1, x2: [B,C,H,W]   out: [B,(2md+1)^2,H,W]   max_disp must be 4 (PWC-Lite,
 * models/pwclite.py:182,208-215: pad_size=4,kernel_size=1,max_displacement=4,
 * stride1=stride2=1).  leaky_slope is used only with USF_FLAG_LEAKY_RELU.
 */
int usf_corr_fwd(const void *x1, const void *x2, void *out, int B, int C, int H, int W,
                 int max_disp, int dtype, unsigned flags, float leaky_slope, void *stream);

/*
 * The same cost volume written into a channel slice of a wider tensor: `out` points at the first of the 81
 * channels inside image 0 and consecutive images are `out_batch_stride` ELEMENTS apart (>= 81*H*W).  With
 * USF_FLAG_LEAKY_RELU this is models/pwclite.py:307-308 + the cost-volume part of the torch.cat at :364-366 in one
 * kernel: the estimator's input buffer [B, 81 + 32 + 2, H, W] receives LeakyReLU(corr) directly.
 * Vectorised stores need out_batch_stride % 4 == 0 (bf16: % 8) next to the usual 16-byte alignment of `out`; other
 * strides take the scalar-store kernels.
 */
int usf_corr_fwd_into(const void *x1, const void *x2, void *out, long long out_batch_stride, int B, int C, int H,
                      int W, int max_disp, int dtype, unsigned flags, float leaky_slope, void *stream);

/*
 * Gradients of the cost volume.  gout: [B,81,H,W].  gx1 / gx2: [B,C,H,W], either
 * may be NULL (not needed).  Both are computed as gathers: no atomics, outputs
 * fully overwritten, bit-identical from run to run.  (Images of <= 256 pixels are
 * processed packed in shared memory; there a NON-FINITE feature value in a column
 * next to the image border can reach a border gradient as 0 * inf.  Finite inputs
 * are exact.)
 */
int usf_corr_bwd(const void *x1, const void *x2, const void *gout, void *gx1, void *gx2, int B,
                 int C, int H, int W, int max_disp, int dtype, unsigned flags, void *stream);

/*
 * Backward warp.  out[b,c,y,x] = bilinear(x[b,c], ix, iy) with
 *   ix = unnormalise(normalise(x + flow[b,0,y,x])), iy likewise with flow[b,1],
 * align_corners=True, computed with the reference's fp32 operation order.
 * x, out: [B,C,H,W] (dtype)   flow: [B,2,H,W] fp32.   H, W >= 2.
 */
int usf_warp_fwd(const void *x, const float *flow, void *out, int B, int C, int H, int W,
                 int pad_mode, int dtype, unsigned flags, void *stream);

/* Bytes of scratch usf_warp_bwd needs when gx != NULL (0 when only gflow is wanted,
 * USF_FLAG_WARP_BWD_ATOMIC is set, or the image is small enough -- H*W <= 1024 -- for
 * the single-kernel path, which keeps its tap lists in shared memory; workspace may
 * then be NULL).  Contents need no initialisation; the pointer passed to usf_warp_bwd
 * must be 16-byte aligned. */
size_t usf_warp_bwd_workspace_bytes(int B, int C, int H, int W, unsigned flags);
/* Leading bytes of that workspace that hold the tap counters (see USF_FLAG_WORKSPACE_CLEAN); 0 when
 * no workspace is needed. */
size_t usf_warp_bwd_workspace_clean_bytes(int B, int C, int H, int W, unsigned flags);

/*
 * Gradients of the backward warp.  gout: [B,C,H,W].  gx: [B,C,H,W] or NULL;
 * gflow: [B,2,H,W] fp32 or NULL.  Both are fully overwritten.
 */
int usf_warp_bwd(const void *x, const float *flow, const void *gout, void *gx, float *gflow, int B,
                 int C, int H, int W, int pad_mode, int dtype, unsigned flags, void *workspace,
                 size_t workspace_bytes, void *stream);

/*
 * Forward-backward consistency mask.  With w = flow_warp(flow21, flow12, pad="zeros"):
 *   occ[b,0,y,x] = |flow12 + w|^2 > scale * (|flow12|^2 + |w|^2) + bias ? 1 : 0
 * flow12, flow21: [B,2,H,W] fp32.  occ: [B,1,H,W] fp32 (the reference returns
 * occ.float()).  H, W >= 2.  flags: USF_FLAG_COORD_RECIPROCAL or 0.  No gradient
 * exists (the reference's result goes through `>`).
 */
int usf_occ_bidirection(const float *flow12, const float *flow21, float *occ, int B, int H, int W,
                        float scale, float bias, unsigned flags, void *stream);

/* Bytes of scratch usf_occ_backward needs (8 per pixel: the fixed-point range map).
 * Contents need no initialisation; the pointer must be 8-byte aligned. */
size_t usf_occ_backward_workspace_bytes(int B, int H, int W);

/*
 * Range-map occlusion mask.  Every source pixel p of flow21 splats the weight 1,
 * split bilinearly, onto the 4 integer neighbours of p + flow21[p] that lie inside
 * the image; occ = clamp(range, 0, 1) < th ? 1 : 0.  The sum is accumulated in
 * 2^-40 fixed point with integer atomics, so the result is bit-identical from run
 * to run.  Non-finite coordinates contribute nothing.
 * flow21: [B,2,H,W] fp32.  occ: [B,1,H,W] fp32.  range_map: [B,1,H,W] fp32 or NULL
 * (the unclamped sum, get_corresponding_map's return value).  H, W <= 2^24.
 */
int usf_occ_backward(const float *flow21, float *occ, float *range_map, int B, int H, int W, float th,
                     void *workspace, size_t workspace_bytes, void *stream);

/*
 * Mask-feature pooling of the "+mf" configurations (models/pwclite.py:317-340, both images):
 *   out[b,c,p] = max over the pixels q of p's segment of feat[b,c,q], clamped at 0 from below unless the segment
 *   covers the whole image (the reference's amax runs over one_hot(seg) * feat, where the pixels outside a segment
 *   contribute zeros) -- what `(one_hot * amax(one_hot * feat))).sum(-1)` computes, without its [B,C,h,w,N]
 *   temporaries.
 * feat, out: [B,C,H,W] (dtype).  seg: [B,H,W] int32 segment ids in [0, num_segments) (ids outside that range belong
 * to no segment: out = 0, no gradient).  argmax: [B,C,num_segments] int32 or NULL (inference): the pixel index that
 * attains each maximum (smallest index on ties), -1 where the clamp at 0 won; usf_segpool_bwd needs it.
 * num_segments <= usf_segpool_max_segments().  Bit-reproducible (max is order independent).
 */
int usf_segpool_max_segments(void);
int usf_segpool_fwd(const void *feat, const int *seg, void *out, int *argmax, int B, int C, int H, int W,
                    int num_segments, int dtype, void *stream);
/*
 * Gradient of the pooling: gfeat[b,c,p] = sum of gout[b,c,.] over p's segment if p == argmax[b,c,seg[p]], else 0.
 * The segment sums are formed in a fixed order (no float atomics): bit-reproducible.  Where several pixels tie for
 * a maximum torch.amax splits the gradient evenly; here the first pixel takes it (measure zero for real features).
 */
int usf_segpool_bwd(const void *gout, const int *seg, const int *argmax, void *gfeat, int B, int C, int H, int W,
                    int num_segments, int dtype, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* UNSAMFLOW_B200_H_ */
