// occ.cu -- the two occlusion masks of the photometric loss ("next" row 3) for sm_100a.
//
// Replaces:
//   utils/warp_utils.py:109-117  get_occu_mask_bidirection: flow_warp(flow21, flow12, pad="zeros") followed by
//       seven elementwise / reduction kernels and a compare            -> occ_bidirection_kernel (one pass)
//   utils/warp_utils.py:26-94,120-126  get_occu_mask_backward: mesh_grid on the CPU + H2D copy, a 4-tap
//       forward splat of weight 1 per source pixel through scatter_add_ (~40 ATen kernels, float atomics on a
//       GPU), clamp and compare                                        -> occ_splat_kernel + occ_threshold_kernel
//
// Both are HBM/L2-bound single passes over [B,2,H,W] flow fields; nothing is staged.
//
// The forward splat accumulates in 64-bit FIXED POINT (2^-40 units, integer red.global.add.u64): integer addition
// is associative, so the range map -- and the mask thresholded from it -- is bit-identical from run to run whatever
// order the taps arrive in, which a float-atomic scatter (what the reference does on a GPU) cannot promise.  Each
// weight is the reference's own fp32 product, quantised with an error below 2^-41; the sum is then exact, i.e. it
// sits inside the fp32 summation noise of the reference's sequential scatter_add_ (~1e-7).
#include "common.cuh"
#include "warp_coord.cuh"

namespace usf {
namespace {

constexpr int kBlock = 256;
constexpr float kFixScale = 1099511627776.0f;          // 2^40
constexpr float kFixInv = 1.0f / 1099511627776.0f;     // 2^-40 (exact)

// flow_warp(flow21, flow12, pad="zeros") and the forward-backward consistency test in one pass.
// Arithmetic order of warp_utils.py:110-117: channel sums of two terms are one rounded add each.
__global__ void __launch_bounds__(kBlock)
occ_bidirection_kernel(const float *__restrict__ flow12, const float *__restrict__ flow21,
                       float *__restrict__ occ, int N, int H, int W, float scale, float bias, int rcp,
                       FastDiv dHW, FastDiv dW) {
    pdl_wait();
    const int p = blockIdx.x * kBlock + threadIdx.x;
    if (p >= N) return;
    const int HW = H * W;
    const int b = dHW.div(p), rem = p - b * HW;
    const int y = dW.div(rem), xo = rem - y * W;
    const float *f12 = flow12 + (size_t)b * 2 * HW;
    const float u = __ldg(f12 + rem), v = __ldg(f12 + HW + rem);
    const Taps t = make_taps(y, xo, u, v, H, W, /*border=*/false, rcp != 0);
    const float nw = t.ex * t.ey, ne = t.wx * t.ey, sw = t.ex * t.wy, se = t.wx * t.wy;
    const float *q0 = flow21 + (size_t)b * 2 * HW + t.o_nw;  // only dereferenced where the tap is inside
    float wv[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const float *q = q0 + (size_t)c * HW;
        float acc = (t.k_nw ? __ldg(q) : 0.f) * nw;
        acc = fmaf(t.k_ne ? __ldg(q + 1) : 0.f, ne, acc);
        acc = fmaf(t.k_sw ? __ldg(q + W) : 0.f, sw, acc);
        acc = fmaf(t.k_se ? __ldg(q + W + 1) : 0.f, se, acc);
        wv[c] = acc;
    }
    const float d0 = __fadd_rn(u, wv[0]), d1 = __fadd_rn(v, wv[1]);
    const float mag = __fadd_rn(__fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v)),
                                __fadd_rn(__fmul_rn(wv[0], wv[0]), __fmul_rn(wv[1], wv[1])));
    const float thresh = __fadd_rn(__fmul_rn(scale, mag), bias);
    const float lhs = __fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1));
    occ[p] = lhs > thresh ? 1.0f : 0.0f;
}

// One axis of get_corresponding_map (warp_utils.py:42-56, 78-85): the floor / ceil neighbour, whether it is
// inside the image (the reference tests `unclamped != clamped`) and its weight 1 - |x - neighbour|.
struct Axis {
    int lo;          // floor neighbour (valid only if lo_ok); the ceil neighbour is lo + 1
    bool lo_ok, hi_ok;
    float w_lo, w_hi;
};
__device__ __forceinline__ Axis splat_axis(int i, float fl, int size) {
    Axis a;
    const float x = __fadd_rn((float)i, fl);
    const float x1 = floorf(x), x0 = __fadd_rn(x1, 1.0f);
    const float mx = (float)(size - 1);
    // NaN / inf coordinates fail every comparison: no tap (the reference's .long() of them is undefined)
    a.lo_ok = (x1 >= 0.f) && (x1 <= mx);
    a.hi_ok = (x0 >= 0.f) && (x0 <= mx);
    a.lo = (int)fminf(fmaxf(x1, -1.0f), mx);
    a.w_lo = __fsub_rn(1.0f, fabsf(__fsub_rn(x, x1)));
    a.w_hi = __fsub_rn(1.0f, fabsf(__fsub_rn(x, x0)));
    return a;
}

__device__ __forceinline__ void splat_add(unsigned long long *plane, int off, bool ok, float w) {
    if (ok && w > 0.f) atomicAdd(plane + off, __float2ull_rn(__fmul_rn(w, kFixScale)));
}

__global__ void __launch_bounds__(kBlock)
occ_splat_kernel(const float *__restrict__ flow21, unsigned long long *__restrict__ acc, int N, int H, int W,
                 FastDiv dHW, FastDiv dW) {
    pdl_wait();
    const int p = blockIdx.x * kBlock + threadIdx.x;
    if (p >= N) return;
    const int HW = H * W;
    const int b = dHW.div(p), rem = p - b * HW;
    const int y = dW.div(rem), xo = rem - y * W;
    const float *f = flow21 + (size_t)b * 2 * HW;
    const Axis ax = splat_axis(xo, __ldg(f + rem), W);
    const Axis ay = splat_axis(y, __ldg(f + HW + rem), H);
    unsigned long long *plane = acc + (size_t)b * HW;
    const int o = ay.lo * W + ax.lo;
    splat_add(plane, o + W + 1, ax.hi_ok && ay.hi_ok, __fmul_rn(ax.w_hi, ay.w_hi));  // ceil, ceil
    splat_add(plane, o + 1, ax.hi_ok && ay.lo_ok, __fmul_rn(ax.w_hi, ay.w_lo));      // ceil x, floor y
    splat_add(plane, o + W, ax.lo_ok && ay.hi_ok, __fmul_rn(ax.w_lo, ay.w_hi));      // floor x, ceil y
    splat_add(plane, o, ax.lo_ok && ay.lo_ok, __fmul_rn(ax.w_lo, ay.w_lo));          // floor, floor
}

// occ = clamp(range, 0, 1) < th (warp_utils.py:124-126); the range map itself is written too when asked for.
__global__ void __launch_bounds__(kBlock)
occ_threshold_kernel(const unsigned long long *__restrict__ acc, float *__restrict__ occ,
                     float *__restrict__ range_map, int N, float th) {
    pdl_wait();
    const int p = blockIdx.x * kBlock + threadIdx.x;
    if (p >= N) return;
    const float r = __fmul_rn(__ull2float_rn(acc[p]), kFixInv);  // one rounding of the exact sum
    if (range_map) range_map[p] = r;
    occ[p] = fminf(fmaxf(r, 0.f), 1.f) < th ? 1.0f : 0.0f;
}

}  // namespace

int occ_bidirection(const float *flow12, const float *flow21, float *occ, int B, int H, int W, float scale,
                    float bias, unsigned flags, cudaStream_t stream) {
    const int N = B * H * W;
    const int blocks = (N + kBlock - 1) / kBlock;
    return note_launch(usf_launch(occ_bidirection_kernel, dim3(blocks), dim3(kBlock), 0, stream, flow12, flow21,
                                  occ, N, H, W, scale, bias, (flags & USF_FLAG_COORD_RECIPROCAL) ? 1 : 0,
                                  make_fastdiv(H * W), make_fastdiv(W)));
}

size_t occ_backward_workspace_bytes(int B, int H, int W) { return (size_t)B * H * W * sizeof(unsigned long long); }

int occ_backward(const float *flow21, float *occ, float *range_map, int B, int H, int W, float th,
                 void *workspace, cudaStream_t stream) {
    const int N = B * H * W;
    const int blocks = (N + kBlock - 1) / kBlock;
    unsigned long long *acc = (unsigned long long *)workspace;
    cudaError_t e = cudaMemsetAsync(acc, 0, (size_t)N * sizeof(unsigned long long), stream);
    if (e != cudaSuccess) { note_cuda_error(e); return USF_ERR_CUDA; }
    int st = note_launch(usf_launch(occ_splat_kernel, dim3(blocks), dim3(kBlock), 0, stream, flow21, acc, N, H, W,
                                    make_fastdiv(H * W), make_fastdiv(W)));
    if (st != USF_OK) return st;
    return note_launch(usf_launch(occ_threshold_kernel, dim3(blocks), dim3(kBlock), 0, stream,
                                  (const unsigned long long *)acc, occ, range_map, N, th));
}

}  // namespace usf
