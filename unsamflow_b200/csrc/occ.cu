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
#include <stdlib.h>

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

__device__ __forceinline__ unsigned long long splat_fix(bool ok, float w) {
    return (ok && w > 0.f) ? __float2ull_rn(__fmul_rn(w, kFixScale)) : 0ull;
}

// MERGE: the x-neighbouring lane of a smooth flow lands in the next cell of the same row, so its (floor x) taps
// are this lane's (ceil x) taps.  The receiving lane (i+1) folds lane i's two ceil-x weights into its own floor-x
// weights before the atomic: two REDs per pixel instead of four.  The fold is an INTEGER add of the same
// fixed-point terms, so the result is bit-identical to the unmerged kernel (USF_OCC_MERGE=0).  Only target
// coordinates are compared, never source adjacency, so row ends and image boundaries need no special case beyond
// the image index; a non-zero folded term implies the giver's tap was inside the image, which makes the
// receiver's address (the same coordinates) valid even when its own taps are not.
// ITER pixels per thread, kBlock apart (lanes stay x-neighbours in every round): all 2*ITER flow loads are issued
// before the first is consumed, and the grid shrinks ITER-fold -- with one pixel per thread the kernel is a chain
// of load latency -> atomics -> exit per warp and spends its time launching CTAs (ncu: 26 us for 1.7 M pixels at
// 12 % of DRAM bandwidth).
template <bool MERGE, int ITER>
__global__ void __launch_bounds__(kBlock)
occ_splat_kernel(const float *__restrict__ flow21, unsigned long long *__restrict__ acc, int N, int H, int W,
                 FastDiv dHW, FastDiv dW) {
    pdl_wait();
    const int HW = H * W;
    const int base = blockIdx.x * (kBlock * ITER) + threadIdx.x;
    float u[ITER], v[ITER];
    int bi[ITER], ri[ITER];
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int p = base + it * kBlock;
        u[it] = v[it] = 0.f;
        bi[it] = -1;
        ri[it] = 0;
        if (p < N) {
            bi[it] = dHW.div(p);
            ri[it] = p - bi[it] * HW;
            const float *f = flow21 + (size_t)bi[it] * 2 * HW;
            u[it] = __ldg(f + ri[it]);
            v[it] = __ldg(f + HW + ri[it]);
        }
    }
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int b = bi[it];
        const bool live = b >= 0;
        int xlo = -1, ylo = -1;
        unsigned long long t_ll = 0, t_hl = 0, t_lh = 0, t_hh = 0;  // (x lo|hi, y lo|hi) in 2^-40 units
        if (live) {
            const int y = dW.div(ri[it]), xo = ri[it] - y * W;
            const Axis ax = splat_axis(xo, u[it], W);
            const Axis ay = splat_axis(y, v[it], H);
            xlo = ax.lo;
            ylo = ay.lo;
            t_hh = splat_fix(ax.hi_ok && ay.hi_ok, __fmul_rn(ax.w_hi, ay.w_hi));  // ceil, ceil
            t_hl = splat_fix(ax.hi_ok && ay.lo_ok, __fmul_rn(ax.w_hi, ay.w_lo));  // ceil x, floor y
            t_lh = splat_fix(ax.lo_ok && ay.hi_ok, __fmul_rn(ax.w_lo, ay.w_hi));  // floor x, ceil y
            t_ll = splat_fix(ax.lo_ok && ay.lo_ok, __fmul_rn(ax.w_lo, ay.w_lo));  // floor, floor
        }
        if (MERGE) {  // every lane of the warp is here in every round (no early exit)
            const unsigned full = 0xffffffffu;
            const int lane = threadIdx.x & 31;
            const int pb = __shfl_up_sync(full, b, 1), px = __shfl_up_sync(full, xlo, 1), py = __shfl_up_sync(full, ylo, 1);
            const unsigned long long p_hl = __shfl_up_sync(full, t_hl, 1), p_hh = __shfl_up_sync(full, t_hh, 1);
            const bool take = lane > 0 && live && pb == b && px + 1 == xlo && py == ylo;
            const bool given = __shfl_down_sync(full, take ? 1 : 0, 1) != 0 && lane < 31;
            if (take) { t_ll += p_hl; t_lh += p_hh; }
            if (given) { t_hl = 0; t_hh = 0; }
        }
        if (live) {
            unsigned long long *plane = acc + (size_t)b * HW;
            const int o = ylo * W + xlo;
            if (t_ll) atomicAdd(plane + o, t_ll);
            if (t_hl) atomicAdd(plane + o + 1, t_hl);
            if (t_lh) atomicAdd(plane + o + W, t_lh);
            if (t_hh) atomicAdd(plane + o + W + 1, t_hh);
        }
    }
}

// USF_OCC_TILE=1 (opt-in): one CTA per 16x64 tile of SOURCE pixels.  Taps are accumulated with shared-memory
// integer atomics in a window of 33x81 targets placed where the tile's centre pixel lands (+-8 pixels of slack
// around the displaced tile), and the window is flushed with one RED per touched target, consecutive lanes on
// consecutive targets: ~1 coalesced global RED per pixel instead of 4 scattered ones (the default kernel is bound
// by the number of RED sectors).  Taps that fall outside the window take the global path.  All
// adds are integer adds of the same fixed-point terms: bit-identical to the other variants.
constexpr int kTH = 16, kTW = 64, kR = 8;
constexpr int kWH = kTH + 2 * kR + 1, kWW = kTW + 2 * kR + 1;

__global__ void __launch_bounds__(kBlock)
occ_splat_tile_kernel(const float *__restrict__ flow21, unsigned long long *__restrict__ acc, int H, int W,
                      int tiles_x, int tiles_y) {
    __shared__ unsigned long long win[kWH * kWW];
    __shared__ int org[2];
    pdl_wait();
    const int HW = H * W;
    int t = blockIdx.x;
    const int tx0 = (t % tiles_x) * kTW;
    t /= tiles_x;
    const int ty0 = (t % tiles_y) * kTH;
    const int b = t / tiles_y;
    const float *f = flow21 + (size_t)b * 2 * HW;
    if (threadIdx.x == 0) {
        const int cx = min(tx0 + kTW / 2, W - 1), cy = min(ty0 + kTH / 2, H - 1);
        const float cu = __ldg(f + cy * W + cx), cv = __ldg(f + HW + cy * W + cx);
        // non-finite or huge centre flow: any window position is correct (it only decides which taps go global)
        const int ou = (fabsf(cu) < 1048576.f) ? __float2int_rn(cu) : 0;
        const int ov = (fabsf(cv) < 1048576.f) ? __float2int_rn(cv) : 0;
        org[0] = tx0 + ou - kR;
        org[1] = ty0 + ov - kR;
    }
    for (int i = threadIdx.x; i < kWH * kWW; i += kBlock) win[i] = 0ull;
    const int lx = threadIdx.x % kTW, ly = threadIdx.x / kTW;  // 64 columns x 4 rows of threads, 4 rows each
    constexpr int RPT = kTH / (kBlock / kTW);
    const int xo = tx0 + lx;
    float u[RPT], v[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        const int y = ty0 + ly + k * (kBlock / kTW);
        u[k] = v[k] = 0.f;
        if (xo < W && y < H) {
            u[k] = __ldg(f + y * W + xo);
            v[k] = __ldg(f + HW + y * W + xo);
        }
    }
    __syncthreads();
    const int wx0 = org[0], wy0 = org[1];
    unsigned long long *plane = acc + (size_t)b * HW;
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        const int y = ty0 + ly + k * (kBlock / kTW);
        if (xo < W && y < H) {
            const Axis ax = splat_axis(xo, u[k], W);
            const Axis ay = splat_axis(y, v[k], H);
            const unsigned long long tap[4] = {
                splat_fix(ax.lo_ok && ay.lo_ok, __fmul_rn(ax.w_lo, ay.w_lo)),
                splat_fix(ax.hi_ok && ay.lo_ok, __fmul_rn(ax.w_hi, ay.w_lo)),
                splat_fix(ax.lo_ok && ay.hi_ok, __fmul_rn(ax.w_lo, ay.w_hi)),
                splat_fix(ax.hi_ok && ay.hi_ok, __fmul_rn(ax.w_hi, ay.w_hi))};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (tap[j]) {  // non-zero implies the target is inside the image
                    const int X = ax.lo + (j & 1), Y = ay.lo + (j >> 1);
                    const int rx = X - wx0, ry = Y - wy0;
                    if ((unsigned)rx < (unsigned)kWW && (unsigned)ry < (unsigned)kWH)
                        atomicAdd(&win[ry * kWW + rx], tap[j]);
                    else
                        atomicAdd(plane + Y * W + X, tap[j]);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kWH * kWW; i += kBlock) {
        const unsigned long long s = win[i];
        if (s) {
            const int ry = i / kWW, rx = i - ry * kWW;
            atomicAdd(plane + (wy0 + ry) * W + (wx0 + rx), s);
        }
    }
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
    static const bool merge = [] {
        const char *e = getenv("USF_OCC_MERGE");
        return !(e && e[0] == '0');
    }();
    static const int iter = [] {  // pixels per thread of the splat: 1 or 4 (A/B switch)
        const char *e = getenv("USF_OCC_ITER");
        return (e && e[0] == '4') ? 4 : 1;
    }();
    static const bool tile = [] {  // opt-in shared-memory window variant
        const char *e = getenv("USF_OCC_TILE");
        return e && e[0] == '1';
    }();
    if (tile) {
        const int tiles_x = (W + kTW - 1) / kTW, tiles_y = (H + kTH - 1) / kTH;
        int st = note_launch(usf_launch(occ_splat_tile_kernel, dim3((unsigned)((size_t)B * tiles_x * tiles_y)), dim3(kBlock), 0,
                                        stream, flow21, acc, H, W, tiles_x, tiles_y));
        if (st != USF_OK) return st;
        return note_launch(usf_launch(occ_threshold_kernel, dim3(blocks), dim3(kBlock), 0, stream,
                                      (const unsigned long long *)acc, occ, range_map, N, th));
    }
    const int sblocks = (N + kBlock * iter - 1) / (kBlock * iter);
    cudaError_t le;
#define USF_SPLAT(M, I) \
    le = usf_launch(occ_splat_kernel<M, I>, dim3(sblocks), dim3(kBlock), 0, stream, flow21, acc, N, H, W, \
                    make_fastdiv(H * W), make_fastdiv(W))
    if (iter == 4) { if (merge) USF_SPLAT(true, 4); else USF_SPLAT(false, 4); }
    else { if (merge) USF_SPLAT(true, 1); else USF_SPLAT(false, 1); }
#undef USF_SPLAT
    int st = note_launch(le);
    if (st != USF_OK) return st;
    return note_launch(usf_launch(occ_threshold_kernel, dim3(blocks), dim3(kBlock), 0, stream,
                                  (const unsigned long long *)acc, occ, range_map, N, th));
}

}  // namespace usf
