// corr_direct.cu -- cost volume fwd/bwd straight from global memory, for what no other kernel takes.
//
// Dispatch order of the correlation (api.cu, corr_fwd.cu, corr_bwd.cu): whole images of <= 256 pixels go
// to corr_small.cu (KITTI 4x13 / 8x26), widths that TMA can address to the tiled kernels, large odd widths
// to the cp.async tiled kernels.  What is left comes here: SMALL shapes (B*H*W <= 4096) above 256 pixels per
// image whose width is not a multiple of 4 (fp32) -- e.g. crops like 9x33 -- and bf16 shapes whose width is
// not a multiple of 8.  Everything is L1/L2 resident at these sizes and tiling through shared memory only
// adds barriers, so these kernels read global memory directly (coalesced along x):
//   fwd:  32 outputs x 8 channel slices per CTA (one warp per slice), partial sums reduced in a fixed
//         order through shared memory: short dependent-load chains, plenty of threads.
//   bwd:  one thread = 4 channels of one pixel x 3 displacement rows, BOTH gradients as gathers.
// Same formulas as corr_fwd.cu / corr_bwd.cu (models/correlation_native.py:13-23).
#include "common.cuh"

namespace usf {
namespace {

// ---- forward: 32 outputs x 8 channel slices per CTA, fixed-order reduction through shared memory ----
constexpr int kFwdOutputs = 32, kFwdSlices = 8, kFwdThreads = kFwdOutputs * kFwdSlices;

template <bool LEAKY, typename T>
__global__ void __launch_bounds__(kFwdThreads)
corr_fwd_direct(const T *__restrict__ x1, const T *__restrict__ x2, T *__restrict__ out, int total,
                int C, int H, int W, float rc, float slope, size_t obs) {
    __shared__ float red[kFwdSlices][kFwdOutputs];
    const int lane = threadIdx.x & 31, cs = threadIdx.x >> 5;  // one warp == one channel slice
    const int idx = blockIdx.x * kFwdOutputs + lane;
    const int HW = H * W;
    float acc = 0.f;
    if (idx < total) {
        const int x = idx % W, y = (idx / W) % H, k = (idx / HW) % kK, b = idx / (HW * kK);
        const int yy = y + k / kD - kMaxDisp, xx = x + k % kD - kMaxDisp;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            const size_t step = (size_t)kFwdSlices * HW;
            const T *p1 = x1 + ((size_t)b * C + cs) * HW + y * W + x;
            const T *p2 = x2 + ((size_t)b * C + cs) * HW + yy * W + xx;
            float a0 = 0.f, a1 = 0.f;
            int c = cs;
            for (; c + 3 * kFwdSlices < C; c += 4 * kFwdSlices, p1 += 4 * step, p2 += 4 * step) {
                const float u0 = ldf(p1), v0 = ldf(p2), u1 = ldf(p1 + step), v1 = ldf(p2 + step);
                const float u2 = ldf(p1 + 2 * step), v2 = ldf(p2 + 2 * step);
                const float u3 = ldf(p1 + 3 * step), v3 = ldf(p2 + 3 * step);
                a0 = fmaf(u0, v0, a0); a1 = fmaf(u1, v1, a1); a0 = fmaf(u2, v2, a0); a1 = fmaf(u3, v3, a1);
            }
            for (; c < C; c += kFwdSlices, p1 += step, p2 += step) a0 = fmaf(ldf(p1), ldf(p2), a0);
            acc = a0 + a1;
        }
    }
    red[cs][lane] = acc;
    __syncthreads();
    if (cs != 0 || idx >= total) return;
    float sum = red[0][lane];
#pragma unroll
    for (int s = 1; s < kFwdSlices; ++s) sum += red[s][lane];
    float v = sum * rc;
    if (LEAKY) v = v < 0.f ? v * slope : v;
    const int ob = idx / (HW * kK);  // image index: the cost volume may be a channel slice of a wider tensor
    stf(out + (size_t)ob * obs + (idx - ob * HW * kK), v);
}

// ---- backward: one thread = 4 channels of one pixel x one third of the displacement rows ----------
// gO[k,y,x] and gO[k,y-dy,x-dx] are loaded once and reused for 4 channels (10 loads per 8 FFMA);
// the 3 row-slices of a pixel live in different warps and are summed through shared memory.
constexpr int kBwdItems = 64, kBwdSlices = 3, kBwdThreads = kBwdItems * kBwdSlices, kCC = 4;

template <bool G1, bool G2, typename T>
__global__ void __launch_bounds__(kBwdThreads)
corr_bwd_direct(const T *__restrict__ x1, const T *__restrict__ x2, const T *__restrict__ gout,
                T *__restrict__ gx1, T *__restrict__ gx2, int items, int C, int H, int W, float rc) {
    __shared__ float red[kBwdSlices][2 * kCC][kBwdItems];
    const int li = threadIdx.x % kBwdItems, ks = threadIdx.x / kBwdItems;
    const int item = blockIdx.x * kBwdItems + li;
    const int HW = H * W, CG = (C + kCC - 1) / kCC;
    float a1[kCC] = {0.f, 0.f, 0.f, 0.f}, a2[kCC] = {0.f, 0.f, 0.f, 0.f};
    int x = 0, y = 0, c0 = 0, b = 0;
    if (item < items) {
        x = item % W; y = (item / W) % H; c0 = ((item / HW) % CG) * kCC; b = item / (HW * CG);
        const T *pg = gout + (size_t)b * kK * HW;
        const T *p1 = x1 + ((size_t)b * C + c0) * HW;
        const T *p2 = x2 + ((size_t)b * C + c0) * HW;
        const int nc = min(kCC, C - c0);
#pragma unroll 1
        for (int dyi = 3 * ks; dyi < 3 * ks + 3; ++dyi) {
            const int dy = dyi - kMaxDisp, yp = y + dy, ym = y - dy;
            const bool ypok = yp >= 0 && yp < H, ymok = ym >= 0 && ym < H;
#pragma unroll
            for (int dxi = 0; dxi < kD; ++dxi) {
                const int dx = dxi - kMaxDisp, k = dyi * kD + dxi;
                if (G1) {
                    const int xp = x + dx;
                    if (ypok && xp >= 0 && xp < W) {
                        const float g = ldf(pg + (size_t)k * HW + y * W + x);
                        const T *q = p2 + yp * W + xp;
#pragma unroll
                        for (int cc = 0; cc < kCC; ++cc)
                            if (cc < nc) a1[cc] = fmaf(g, ldf(q + (size_t)cc * HW), a1[cc]);
                    }
                }
                if (G2) {
                    const int xm = x - dx;
                    if (ymok && xm >= 0 && xm < W) {
                        const float g = ldf(pg + (size_t)k * HW + ym * W + xm);
                        const T *q = p1 + ym * W + xm;
#pragma unroll
                        for (int cc = 0; cc < kCC; ++cc)
                            if (cc < nc) a2[cc] = fmaf(g, ldf(q + (size_t)cc * HW), a2[cc]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int cc = 0; cc < kCC; ++cc) {
        red[ks][cc][li] = a1[cc];
        red[ks][kCC + cc][li] = a2[cc];
    }
    __syncthreads();
    if (ks != 0 || item >= items) return;
    const int nc = min(kCC, C - c0);
    const size_t o = ((size_t)b * C + c0) * HW + y * W + x;
#pragma unroll
    for (int cc = 0; cc < kCC; ++cc) {
        if (cc >= nc) break;
        if (G1) stf(gx1 + o + (size_t)cc * HW, ((red[0][cc][li] + red[1][cc][li]) + red[2][cc][li]) * rc);
        if (G2) stf(gx2 + o + (size_t)cc * HW, ((red[0][kCC + cc][li] + red[1][kCC + cc][li]) + red[2][kCC + cc][li]) * rc);
    }
}

}  // namespace

template <typename T>
static int fwd_direct(const T *x1, const T *x2, T *out, int B, int C, int H, int W, unsigned flags, float slope,
                      cudaStream_t stream) {
    const int total = B * kK * H * W;
    const int blocks = (total + kFwdOutputs - 1) / kFwdOutputs;
    const float rc = 1.0f / (float)C;
    if (flags & USF_FLAG_LEAKY_RELU)
        corr_fwd_direct<true, T><<<blocks, kFwdThreads, 0, stream>>>(x1, x2, out, total, C, H, W, rc, slope, corr_out_bstride(H, W));
    else
        corr_fwd_direct<false, T><<<blocks, kFwdThreads, 0, stream>>>(x1, x2, out, total, C, H, W, rc, slope, corr_out_bstride(H, W));
    return note_launch(cudaGetLastError());
}

template <typename T>
static int bwd_direct(const T *x1, const T *x2, const T *gout, T *gx1, T *gx2, int B, int C, int H, int W,
                      cudaStream_t stream) {
    const int items = B * ((C + kCC - 1) / kCC) * H * W;
    const int blocks = (items + kBwdItems - 1) / kBwdItems;
    const float rc = 1.0f / (float)C;
    if (gx1 && gx2)
        corr_bwd_direct<true, true, T><<<blocks, kBwdThreads, 0, stream>>>(x1, x2, gout, gx1, gx2, items, C, H, W, rc);
    else if (gx1)
        corr_bwd_direct<true, false, T><<<blocks, kBwdThreads, 0, stream>>>(x1, x2, gout, gx1, gx2, items, C, H, W, rc);
    else
        corr_bwd_direct<false, true, T><<<blocks, kBwdThreads, 0, stream>>>(x1, x2, gout, gx1, gx2, items, C, H, W, rc);
    return note_launch(cudaGetLastError());
}

int corr_fwd_direct_f32(const float *x1, const float *x2, float *out, int B, int C, int H, int W, unsigned flags,
                        float slope, cudaStream_t stream) {
    return fwd_direct<float>(x1, x2, out, B, C, H, W, flags, slope, stream);
}
int corr_bwd_direct_f32(const float *x1, const float *x2, const float *gout, float *gx1, float *gx2, int B, int C,
                        int H, int W, cudaStream_t stream) {
    return bwd_direct<float>(x1, x2, gout, gx1, gx2, B, C, H, W, stream);
}
// bf16 storage, fp32 accumulate: the fallback behind corr_small.cu and the tiled bf16 TMA kernels
// (corr_fwd_bf16.cu / corr_bwd_bf16.cu, W % 8 == 0); correct for every shape, slow for large ones
int corr_fwd_bf16(const void *x1, const void *x2, void *out, int B, int C, int H, int W, unsigned flags, float slope,
                  cudaStream_t stream) {
    return fwd_direct<bf16>((const bf16 *)x1, (const bf16 *)x2, (bf16 *)out, B, C, H, W, flags, slope, stream);
}
int corr_bwd_bf16(const void *x1, const void *x2, const void *gout, void *gx1, void *gx2, int B, int C, int H, int W,
                  cudaStream_t stream) {
    return bwd_direct<bf16>((const bf16 *)x1, (const bf16 *)x2, (const bf16 *)gout, (bf16 *)gx1, (bf16 *)gx2, B, C, H,
                            W, stream);
}

}  // namespace usf
