// corr_direct.cu -- cost volume fwd/bwd for TINY pyramid levels (a few thousand pixels or less).
//
// The coarsest PWC-Lite levels (KITTI 4x13 and 8x26, models/pwclite.py:295 levels 0-1) hold
// 52..208 pixels per image: the whole problem fits in L1/L2, has odd widths (no TMA, no 16-byte
// vectors) and is pure launch/latency.  Tiling through shared memory only adds barriers there,
// so these kernels read global memory directly (coalesced along x, everything L1/L2 resident):
//   fwd:  one thread per output element (b,k,y,x), serial sum over channels, 8-way unrolled.
//   bwd:  one thread per input element (b,c,y,x) computes BOTH gradients as 81-term gathers.
// Same formulas as corr_fwd.cu / corr_bwd.cu (models/correlation_native.py:13-23).
#include "common.cuh"

namespace usf {
namespace {

constexpr int kThreads = 256;

template <bool LEAKY>
__global__ void __launch_bounds__(kThreads)
corr_fwd_direct(const float *__restrict__ x1, const float *__restrict__ x2, float *__restrict__ out, int total,
                int C, int H, int W, float rc, float slope) {
    const int idx = blockIdx.x * kThreads + threadIdx.x;
    if (idx >= total) return;
    const int HW = H * W;
    const int x = idx % W, y = (idx / W) % H, k = (idx / HW) % kK, b = idx / (HW * kK);
    const int dy = k / kD - kMaxDisp, dx = k % kD - kMaxDisp;
    const int yy = y + dy, xx = x + dx;
    float acc = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const float *p1 = x1 + (size_t)b * C * HW + y * W + x;
        const float *p2 = x2 + (size_t)b * C * HW + yy * W + xx;
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        int c = 0;
        for (; c + 8 <= C; c += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u & 3] = fmaf(__ldg(p1 + (size_t)(c + u) * HW), __ldg(p2 + (size_t)(c + u) * HW), a[u & 3]);
        }
        for (; c < C; ++c) a[0] = fmaf(__ldg(p1 + (size_t)c * HW), __ldg(p2 + (size_t)c * HW), a[0]);
        acc = (a[0] + a[1]) + (a[2] + a[3]);
    }
    float v = acc * rc;
    if (LEAKY) v = v < 0.f ? v * slope : v;
    out[idx] = v;
}

template <bool G1, bool G2>
__global__ void __launch_bounds__(kThreads)
corr_bwd_direct(const float *__restrict__ x1, const float *__restrict__ x2, const float *__restrict__ gout,
                float *__restrict__ gx1, float *__restrict__ gx2, int total, int C, int H, int W, float rc) {
    const int idx = blockIdx.x * kThreads + threadIdx.x;
    if (idx >= total) return;
    const int HW = H * W;
    const int x = idx % W, y = (idx / W) % H, c = (idx / HW) % C, b = idx / (HW * C);
    const float *pg = gout + (size_t)b * kK * HW;
    const float *p1 = x1 + ((size_t)b * C + c) * HW;
    const float *p2 = x2 + ((size_t)b * C + c) * HW;
    float a1 = 0.f, a2 = 0.f;
#pragma unroll 1
    for (int dyi = 0; dyi < kD; ++dyi) {
        const int dy = dyi - kMaxDisp;
        const int yp = y + dy, ym = y - dy;
#pragma unroll
        for (int dxi = 0; dxi < kD; ++dxi) {
            const int dx = dxi - kMaxDisp, k = dyi * kD + dxi;
            if (G1) {
                const int xp = x + dx;
                if (yp >= 0 && yp < H && xp >= 0 && xp < W)
                    a1 = fmaf(__ldg(pg + (size_t)k * HW + y * W + x), __ldg(p2 + yp * W + xp), a1);
            }
            if (G2) {
                const int xm = x - dx;
                if (ym >= 0 && ym < H && xm >= 0 && xm < W)
                    a2 = fmaf(__ldg(pg + (size_t)k * HW + ym * W + xm), __ldg(p1 + ym * W + xm), a2);
            }
        }
    }
    if (G1) gx1[idx] = a1 * rc;
    if (G2) gx2[idx] = a2 * rc;
}

}  // namespace

int corr_fwd_direct_f32(const float *x1, const float *x2, float *out, int B, int C, int H, int W, unsigned flags,
                        float slope, cudaStream_t stream) {
    const int total = B * kK * H * W;
    const int blocks = (total + kThreads - 1) / kThreads;
    const float rc = 1.0f / (float)C;
    if (flags & USF_FLAG_LEAKY_RELU)
        corr_fwd_direct<true><<<blocks, kThreads, 0, stream>>>(x1, x2, out, total, C, H, W, rc, slope);
    else
        corr_fwd_direct<false><<<blocks, kThreads, 0, stream>>>(x1, x2, out, total, C, H, W, rc, slope);
    return note_launch(cudaGetLastError());
}

int corr_bwd_direct_f32(const float *x1, const float *x2, const float *gout, float *gx1, float *gx2, int B, int C,
                        int H, int W, cudaStream_t stream) {
    const int total = B * C * H * W;
    const int blocks = (total + kThreads - 1) / kThreads;
    const float rc = 1.0f / (float)C;
    if (gx1 && gx2)
        corr_bwd_direct<true, true><<<blocks, kThreads, 0, stream>>>(x1, x2, gout, gx1, gx2, total, C, H, W, rc);
    else if (gx1)
        corr_bwd_direct<true, false><<<blocks, kThreads, 0, stream>>>(x1, x2, gout, gx1, gx2, total, C, H, W, rc);
    else
        corr_bwd_direct<false, true><<<blocks, kThreads, 0, stream>>>(x1, x2, gout, gx1, gx2, total, C, H, W, rc);
    return note_launch(cudaGetLastError());
}

}  // namespace usf
