// corr_bwd.cu -- cost-volume backward for sm_100a (fp32 storage, fp32 FFMA), atomic-free.
//
// Replaces: correlation_cuda.backward (models/correlation_package/correlation_cuda.cc:88-165 ->
// channels_first x2 + per-batch-item loops of correlation_backward_input1 / _input2,
// correlation_cuda_kernel.cu:116-300, :458-513), i.e. autograd of models/correlation_native.py:13-23:
//     gx1[b,c,y,x] = (1/C) sum_{dy,dx} gO[b,k(dy,dx),y,x]       * x2[b,c,y+dy,x+dx]
//     gx2[b,c,y,x] = (1/C) sum_{dy,dx} gO[b,k(dy,dx),y-dy,x-dx] * x1[b,c,y-dy,x-dx]
// With e = -d the second sum reads  sum_e gO[b,k(-e),y+ey,x+ex] * x1[b,c,y+ey,x+ex]  and
// k(-e) = 80 - k(e): both gradients are the SAME gather  out[c,p] = sum_e G[e,p] * X[c,p+e],
// they only differ in how the G tile is fetched (gx2: channel 80-k, shifted by e, zero outside).
//
// One kernel template does both: a CTA owns an 8x32 pixel tile and 32 channels (4 warps x 8
// channels), stages the X halo tile (16x40 per channel) once and streams G one dy at a time
// through a 2-stage cp.async ring.  A thread owns 8 pixels x 8 channels (64 accumulators) and
// keeps the 9x8 G values of the current dy in registers: per (dy, channel) it issues 4 LDS.128
// for 72 FFMA.  Every output element is written exactly once: no atomics, no zero-fill, one
// launch per gradient for the whole batch (the reference launches 2*B kernels).
// When W % 4 == 0 the tiles are fetched by TMA (hardware zero-fill does the bounds work): the X halo
// tile as four 8-channel boxes, the G stage as one 9-channel box (see the TMA section below for how
// the shifted G tile of gx2 is handled).  One producer lane feeds 4 consumer warps through
// full/empty mbarriers (no __syncthreads in the loop).
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"

namespace usf {
namespace {

constexpr int TH = 8, TW = 32;
constexpr int CG = 32;                    // channels per CTA
constexpr int CT = 8;                     // channels per thread
constexpr int NTHREADS = 32 * (CG / CT);  // 128
constexpr int XR = TH + 2 * kMaxDisp;     // 16 halo rows
constexpr int XW = TW + 2 * kMaxDisp;     // 40 halo cols
constexpr int XS = 44;                    // halo row stride (11 chunks)
constexpr int GS = 36;                    // G tile row stride (9 chunks)
constexpr int X_FLOATS = CG * XR * XS;    // 22528
constexpr int G_STAGE = kD * TH * GS;     // 2592
constexpr int SMEM_BYTES = (X_FLOATS + 2 * G_STAGE) * 4;  // 110848

template <bool VEC>
__device__ __forceinline__ void load_x_tile(float *__restrict__ xs, const float *__restrict__ xb, int c0,
                                            int C, int H, int W, int y0, int x0, int tid) {
    constexpr int CHUNKS = CG * XR * (XW / 4);  // 5120
    for (int i = tid; i < CHUNKS; i += NTHREADS) {
        const int c = i / (XR * 10);
        const int rem = i - c * (XR * 10);
        const int r = rem / 10, ch = rem - r * 10;
        const int cc = c0 + c, gy = y0 - kMaxDisp + r, gx = x0 - kMaxDisp + 4 * ch;
        float *dst = xs + (c * XR + r) * XS + 4 * ch;
        const bool rowok = (cc < C) && (gy >= 0) && (gy < H);
        const float *src = xb + ((size_t)cc * H + gy) * W + gx;
        if (VEC) {
            const bool ok = rowok && (gx >= 0) && (gx < W);
            cp_async16(dst, ok ? src : xb, ok ? 16 : 0);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool ok = rowok && (gx + e >= 0) && (gx + e < W);
                cp_async4(dst + e, ok ? src + e : xb, ok ? 4 : 0);
            }
        }
    }
}

// G tile of one dy (9 dx): gs[dxi][r][col].
//   gx1: G[dxi][r][col] = gO[k = dyi*9+dxi][y0+r][x0+col]
//   gx2: G[dxi][r][col] = gO[80-k][y0+r+ey][x0+col+ex],  ey = dyi-4, ex = dxi-4, zero outside
template <bool GX2, bool VEC>
__device__ __forceinline__ void load_g_stage(float *__restrict__ gs, const float *__restrict__ gob, int dyi,
                                             int H, int W, int y0, int x0, int tid) {
    constexpr int CHUNKS = kD * TH * (TW / 4);  // 576
    for (int i = tid; i < CHUNKS; i += NTHREADS) {
        const int dxi = i >> 6, r = (i >> 3) & 7, ch = i & 7;
        const int k = dyi * kD + dxi;
        float *dst = gs + (dxi * TH + r) * GS + 4 * ch;
        if (!GX2) {
            const int gy = y0 + r, gx = x0 + 4 * ch;
            const float *src = gob + ((size_t)k * H + gy) * W + gx;
            if (VEC) {
                const bool ok = (gy < H) && (gx < W);
                cp_async16(dst, ok ? src : gob, ok ? 16 : 0);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const bool ok = (gy < H) && (gx + e < W);
                    cp_async4(dst + e, ok ? src + e : gob, ok ? 4 : 0);
                }
            }
        } else {
            const int ex = dxi - kMaxDisp;
            const int gy = y0 + r + dyi - kMaxDisp, gx = x0 + 4 * ch + ex;
            const bool rowok = (gy >= 0) && (gy < H);
            const float *src = gob + ((size_t)(kK - 1 - k) * H + gy) * W + gx;
            if (VEC && (ex & 3) == 0) {
                const bool ok = rowok && (gx >= 0) && (gx < W);
                cp_async16(dst, ok ? src : gob, ok ? 16 : 0);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const bool ok = rowok && (gx + e >= 0) && (gx + e < W);
                    cp_async4(dst + e, ok ? src + e : gob, ok ? 4 : 0);
                }
            }
        }
    }
}

// grid = (tiles_x, tiles_y, B * ngroups)
template <bool GX2, bool VEC>
__global__ void __launch_bounds__(NTHREADS, 2)
corr_bwd_gather(const float *__restrict__ xin, const float *__restrict__ gout, float *__restrict__ gx, int C,
                int H, int W, int ngroups, float rc) {
    extern __shared__ __align__(16) float smem[];
    float *xs = smem;
    float *gsm = smem + X_FLOATS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = lane >> 2, q = lane & 3;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int b = blockIdx.z / ngroups, cg = blockIdx.z - b * ngroups;
    const int c0 = cg * CG;
    const float *xb = xin + (size_t)b * C * H * W;
    const float *gob = gout + (size_t)b * kK * H * W;

    load_x_tile<VEC>(xs, xb, c0, C, H, W, y0, x0, tid);
    load_g_stage<GX2, VEC>(gsm, gob, 0, H, W, y0, x0, tid);
    cp_async_commit();
    load_g_stage<GX2, VEC>(gsm + G_STAGE, gob, 1, H, W, y0, x0, tid);
    cp_async_commit();

    float acc[CT][8];
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c][i] = 0.f;

    const float *xw = xs + (warp * CT * XR + r) * XS + 8 * q;
#pragma unroll 1
    for (int dyi = 0; dyi < kD; ++dyi) {
        cp_async_wait<1>();
        __syncthreads();
        const float *gs = gsm + (dyi & 1) * G_STAGE + r * GS + 8 * q;
        float g[kD][8];
#pragma unroll
        for (int dx = 0; dx < kD; ++dx) {
            const float4 g0 = lds128(gs + dx * TH * GS), g1 = lds128(gs + dx * TH * GS + 4);
            g[dx][0] = g0.x; g[dx][1] = g0.y; g[dx][2] = g0.z; g[dx][3] = g0.w;
            g[dx][4] = g1.x; g[dx][5] = g1.y; g[dx][6] = g1.z; g[dx][7] = g1.w;
        }
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const float *p = xw + (c * XR + dyi) * XS;
            const float4 w0 = lds128(p), w1 = lds128(p + 4), w2 = lds128(p + 8), w3 = lds128(p + 12);
            const float w[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w,
                                 w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
            for (int dx = 0; dx < kD; ++dx)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[c][i] = fmaf(g[dx][i], w[dx + i], acc[c][i]);
        }
        __syncthreads();  // stage (dyi & 1) fully consumed
        if (dyi + 2 < kD) load_g_stage<GX2, VEC>(gsm + (dyi & 1) * G_STAGE, gob, dyi + 2, H, W, y0, x0, tid);
        cp_async_commit();
    }

    const int gy = y0 + r, gxx = x0 + 8 * q;
    if (gy >= H) return;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        const int cc = c0 + warp * CT + c;
        if (cc >= C) break;
        float *o = gx + (((size_t)b * C + cc) * H + gy) * W + gxx;
        if (VEC) {
            if (gxx < W)
                *reinterpret_cast<float4 *>(o) =
                    make_float4(acc[c][0] * rc, acc[c][1] * rc, acc[c][2] * rc, acc[c][3] * rc);
            if (gxx + 4 < W)
                *reinterpret_cast<float4 *>(o + 4) =
                    make_float4(acc[c][4] * rc, acc[c][5] * rc, acc[c][6] * rc, acc[c][7] * rc);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (gxx + i < W) o[i] = acc[c][i] * rc;
        }
    }
}


// ===========================================================================================
// TMA variant (W % 4 == 0): same consumer math, loads by one producer lane + mbarriers.
//
// TMA constraint found the hard way (sm_100a, driver 580): the innermost box coordinate must be a
// multiple of 16 bytes -- a box starting at x0+ex with ex % 4 != 0 raises "illegal instruction".
// So the SHIFTED G tile of gx2 is fetched as ONE aligned halo box [9 ch][8 rows][44 cols] at
// (x0-4, y0+ey, channel (8-dyi)*9): slot j holds gO channel (8-dyi)*9+j, displacement dxi needs
// channel 80-k = slot 8-dxi, and its 8 values start at column 8q+dxi of that row.  dxi is a
// compile-time constant in the unrolled consumer loop, so the unaligned window is read as 2-3
// aligned LDS.128 and selected with static register indices.
// ===========================================================================================
// channel after which a consumer warp releases the G stage (fence + arrive): early enough for the producer to
// refill ahead, late enough that the fence finds the stage's loads already complete (measured: after channel
// 0 -> 47.6 us, 1 -> 46.0, 3 -> 46.1 at 8x32x64x208; 15.0 / 14.7 / 14.2 us in the 4-channel form at 8x96x16x52)
constexpr int kRelAt = 3;
constexpr int NCW = CG / CT;                                 // 4 consumer warps
constexpr int T_THREADS = 32 * (NCW + 1);                    // + 1 producer warp
constexpr int NBAR = NCW + 4;                                // xfull[4], gfull[2], gempty[2]

// CTT = channels per thread: 8 (CTA = 32 channels, 2 CTAs/SM) when the grid fills the chip, 4 (CTA = 16
// channels, 3 CTAs/SM, twice the CTAs, half the dependent work per warp) when it would not.
template <bool GX2, int CTT = CT> struct GTile {
    static constexpr int kStride = GX2 ? XS : GS;            // 44 (halo box) : 36
    static constexpr int kBytes = kD * TH * kStride * 4;     // 12672 : 10368
    static constexpr int kXW = CTT * XR * XS * 4;            // one warp's halo box: 22528 (8 ch) / 11264 (4 ch)
    static constexpr int kSmem = NCW * kXW + 2 * kBytes + NBAR * 8;  // barriers live at the end
};

// eight G values of displacement dxi for this thread's pixels (row base pointer `row`)
template <bool GX2, int DXI>
__device__ __forceinline__ void load_g(float (&g)[8], const float *__restrict__ stage, int r, int q) {
    constexpr int S = GTile<GX2>::kStride;
    if (!GX2) {
        const float *p = stage + (DXI * TH + r) * S + 8 * q;
        const float4 a = lds128(p), b = lds128(p + 4);
        g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
    } else {
        constexpr int A = 4 * (DXI / 4), O = DXI - A;
        const float *p = stage + ((kD - 1 - DXI) * TH + r) * S + 8 * q + A;
        const float4 a = lds128(p), b = lds128(p + 4);
        if (O == 0) {
            g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
        } else {
            const float4 c = lds128(p + 8);
            const float v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = v[O + i];
        }
    }
}

template <bool GX2, int DXI>
struct GLoader {
    static __device__ __forceinline__ void run(float (&g)[kD][8], const float *stage, int r, int q) {
        load_g<GX2, DXI>(g[DXI], stage, r, q);
        GLoader<GX2, DXI + 1>::run(g, stage, r, q);
    }
};
template <bool GX2>
struct GLoader<GX2, kD> {
    static __device__ __forceinline__ void run(float (&)[kD][8], const float *, int, int) {}
};

// mx: box [8ch][16][44] over the feature map.  mg over gout: gx1 box [9ch][8][36], gx2 box [9ch][8][44].
// grid = (tiles_x, tiles_y, B * ngroups)
template <bool GX2, int CTT>
__device__ __forceinline__ void bwd_tile(unsigned char *smem_raw, const CUtensorMap &mx, const CUtensorMap &mg,
                                         float *__restrict__ gx, int C, int H, int W, int ngroups, int zidx,
                                         int tx, int ty, float rc) {
    constexpr int GB = GTile<GX2, CTT>::kBytes;
    constexpr int XW_BYTES = GTile<GX2, CTT>::kXW;
    constexpr int CG = NCW * CTT;  // channels per CTA (shadows the 32 of the cp.async kernels)
    constexpr int CT = CTT;
    unsigned char *gbase = smem_raw + NCW * XW_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(gbase + 2 * GB);
    uint64_t *xfull = bars, *gfull = bars + NCW, *gempty = bars + NCW + 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int x0 = tx * TW, y0 = ty * TH;
    const int b = zidx / ngroups, cg = zidx - b * ngroups;
    const int c0 = cg * CG;
    if (tid == 0) {
#pragma unroll
        for (int w = 0; w < NCW; ++w) mbar_init(&xfull[w], 1);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            mbar_init(&gfull[s], 1);
            mbar_init(&gempty[s], NCW);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == NCW) {
        if (lane == 0) {
            tma_prefetch_desc(&mx);
            tma_prefetch_desc(&mg);
#pragma unroll 1
            for (int dyi = 0; dyi < kD; ++dyi) {
                const int s = dyi & 1;
                if (dyi >= 2) mbar_wait(&gempty[s], ((dyi >> 1) - 1) & 1);
                mbar_expect_tx(&gfull[s], GB);
                if (!GX2)
                    tma_load_4d(gbase + s * GB, &mg, &gfull[s], x0, y0, dyi * kD, b);
                else
                    tma_load_4d(gbase + s * GB, &mg, &gfull[s], x0 - kMaxDisp, y0 + dyi - kMaxDisp,
                                (kD - 1 - dyi) * kD, b);
                if (dyi == 0) {  // the X halo tiles go out right after the first G stage
#pragma unroll 1
                    for (int w = 0; w < NCW; ++w) {
                        mbar_expect_tx(&xfull[w], XW_BYTES);
                        tma_load_4d(smem_raw + w * XW_BYTES, &mx, &xfull[w], x0 - kMaxDisp, y0 - kMaxDisp,
                                    c0 + w * CT, b);
                    }
                }
            }
        }
        return;
    }

    const int r = lane >> 2, q = lane & 3;
    float acc[CT][8];
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c][i] = 0.f;

    const float *xw = reinterpret_cast<const float *>(smem_raw + warp * XW_BYTES) + r * XS + 8 * q;
    mbar_wait(&xfull[warp], 0);
#pragma unroll 1
    for (int dyi = 0; dyi < kD; ++dyi) {
        const int s = dyi & 1;
        mbar_wait(&gfull[s], (dyi >> 1) & 1);
        float g[kD][8];
        GLoader<GX2, 0>::run(g, reinterpret_cast<const float *>(gbase + s * GB), r, q);
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const float *p = xw + (c * XR + dyi) * XS;
            const float4 w0 = lds128(p), w1 = lds128(p + 4), w2 = lds128(p + 8), w3 = lds128(p + 12);
            const float w[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w,
                                 w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
            for (int dx = 0; dx < kD; ++dx)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[c][i] = fmaf(g[dx][i], w[dx + i], acc[c][i]);
            if (c == kRelAt) {
                // Release the G stage after a few channels' FFMAs (every channel uses every g value) and behind
                // a fence: an arrive that merely follows the LDS instructions does not wait for their data, and
                // the producer's next TMA write then races with loads still in flight (seen as rare, large
                // errors in gx2 at 8x32x64x208; see release_stage_reads in common.cuh).
                release_stage_reads();
                if (lane == 0) mbar_arrive(&gempty[s]);
            }
        }
    }

    const int gy = y0 + r, gxx = x0 + 8 * q;
    if (gy >= H) return;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        const int cc = c0 + warp * CT + c;
        if (cc >= C) break;
        float *o = gx + (((size_t)b * C + cc) * H + gy) * W + gxx;
        if (gxx < W)
            *reinterpret_cast<float4 *>(o) =
                make_float4(acc[c][0] * rc, acc[c][1] * rc, acc[c][2] * rc, acc[c][3] * rc);
        if (gxx + 4 < W)
            *reinterpret_cast<float4 *>(o + 4) =
                make_float4(acc[c][4] * rc, acc[c][5] * rc, acc[c][6] * rc, acc[c][7] * rc);
    }
}

// ONE launch computes both gradients: blockIdx.z selects (gradient, batch item, channel group).  The
// gx1 and gx2 halves of the grid are independent, so the fused grid packs the SMs better than two
// back-to-back launches (twice the CTAs per wave tail) and saves a launch.
// z layout: [0, nz) -> first requested gradient, [nz, 2 nz) -> gx2 when both are requested.
template <int CTT>
__global__ void __launch_bounds__(T_THREADS, CTT == 8 ? 2 : 3)
corr_bwd_gather_tma(const __grid_constant__ CUtensorMap mx1, const __grid_constant__ CUtensorMap mg1,
                    const __grid_constant__ CUtensorMap mx2, const __grid_constant__ CUtensorMap mg2,
                    float *__restrict__ gx1, float *__restrict__ gx2, int C, int H, int W, int ngroups, int nz,
                    float rc) {
    pdl_wait();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    int z = blockIdx.z;
    const bool second = (gx1 == nullptr) || z >= nz;
    if (z >= nz) z -= nz;
    if (second)
        bwd_tile<true, CTT>(smem_raw, mx2, mg2, gx2, C, H, W, ngroups, z, blockIdx.x, blockIdx.y, rc);
    else
        bwd_tile<false, CTT>(smem_raw, mx1, mg1, gx1, C, H, W, ngroups, z, blockIdx.x, blockIdx.y, rc);
}

// -------------------------------------------------------------------------------------------
// 16x16 tile geometry for gx1.  Widths like 208 = 6.5 x 32 leave the last 8x32 tile column half empty; at
// 8x32x64x208 that made 896 CTAs = 3.03 waves of the 296 resident ones, i.e. a fourth round for 8 CTAs.
// 16x16 tiles cover such a level exactly.  Same thread tile (8 pixels x 8 channels), lanes mapped as
// (row = lane & 15, strip = lane >> 4) so that a quarter-warp reads 8 different rows: strides of 7 (X halo,
// 24x28) and 5 (G tile, 16x20) 16-byte chunks are conflict-free.  Only gx1 can take it: the gx2 halo stage
// would need 2 x 16 KB more than two CTAs per SM allow, so a mixed launch runs gx1 on 16x16 and gx2 on 8x32
// tiles (864 CTAs, 3 rounds).
// -------------------------------------------------------------------------------------------
constexpr int SQ = 16;                        // tile edge
constexpr int XR2 = SQ + 2 * kMaxDisp;        // 24 halo rows
constexpr int XS2 = 28;                       // halo row stride (24 columns used)
constexpr int GS2 = 20;                       // G tile row stride (16 columns used)
constexpr int XW2_BYTES = CT * XR2 * XS2 * 4; // 21504
constexpr int G2_BYTES = kD * SQ * GS2 * 4;   // 11520
constexpr int SQ_SMEM = NCW * XW2_BYTES + 2 * G2_BYTES + NBAR * 8;

__device__ __forceinline__ void bwd_tile_sq(unsigned char *smem_raw, const CUtensorMap &mx, const CUtensorMap &mg,
                                            float *__restrict__ gx, int C, int H, int W, int ngroups, int zidx,
                                            int tx, int ty, float rc) {
    unsigned char *gbase = smem_raw + NCW * XW2_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(gbase + 2 * G2_BYTES);
    uint64_t *xfull = bars, *gfull = bars + NCW, *gempty = bars + NCW + 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int x0 = tx * SQ, y0 = ty * SQ;
    const int b = zidx / ngroups, cg = zidx - b * ngroups;
    const int c0 = cg * CG;
    if (tid == 0) {
#pragma unroll
        for (int w = 0; w < NCW; ++w) mbar_init(&xfull[w], 1);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            mbar_init(&gfull[s], 1);
            mbar_init(&gempty[s], NCW);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == NCW) {
        if (lane == 0) {
            tma_prefetch_desc(&mx);
            tma_prefetch_desc(&mg);
#pragma unroll 1
            for (int dyi = 0; dyi < kD; ++dyi) {
                const int s = dyi & 1;
                if (dyi >= 2) mbar_wait(&gempty[s], ((dyi >> 1) - 1) & 1);
                mbar_expect_tx(&gfull[s], G2_BYTES);
                tma_load_4d(gbase + s * G2_BYTES, &mg, &gfull[s], x0, y0, dyi * kD, b);
                if (dyi == 0) {
#pragma unroll 1
                    for (int w = 0; w < NCW; ++w) {
                        mbar_expect_tx(&xfull[w], XW2_BYTES);
                        tma_load_4d(smem_raw + w * XW2_BYTES, &mx, &xfull[w], x0 - kMaxDisp, y0 - kMaxDisp,
                                    c0 + w * CT, b);
                    }
                }
            }
        }
        return;
    }

    const int r = lane & 15, q = lane >> 4;
    float acc[CT][8];
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c][i] = 0.f;

    const float *xw = reinterpret_cast<const float *>(smem_raw + warp * XW2_BYTES) + r * XS2 + 8 * q;
    mbar_wait(&xfull[warp], 0);
#pragma unroll 1
    for (int dyi = 0; dyi < kD; ++dyi) {
        const int s = dyi & 1;
        mbar_wait(&gfull[s], (dyi >> 1) & 1);
        const float *gs = reinterpret_cast<const float *>(gbase + s * G2_BYTES) + r * GS2 + 8 * q;
        float g[kD][8];
#pragma unroll
        for (int dx = 0; dx < kD; ++dx) {
            const float4 g0 = lds128(gs + dx * SQ * GS2), g1 = lds128(gs + dx * SQ * GS2 + 4);
            g[dx][0] = g0.x; g[dx][1] = g0.y; g[dx][2] = g0.z; g[dx][3] = g0.w;
            g[dx][4] = g1.x; g[dx][5] = g1.y; g[dx][6] = g1.z; g[dx][7] = g1.w;
        }
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const float *p = xw + (c * XR2 + dyi) * XS2;
            const float4 w0 = lds128(p), w1 = lds128(p + 4), w2 = lds128(p + 8), w3 = lds128(p + 12);
            const float w[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w,
                                 w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
            for (int dx = 0; dx < kD; ++dx)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[c][i] = fmaf(g[dx][i], w[dx + i], acc[c][i]);
            if (c == kRelAt) {  // release the G stage only after its values have been consumed (see bwd_tile)
                release_stage_reads();
                if (lane == 0) mbar_arrive(&gempty[s]);
            }
        }
    }

    const int gy = y0 + r, gxx = x0 + 8 * q;
    if (gy >= H) return;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        const int cc = c0 + warp * CT + c;
        if (cc >= C) break;
        float *o = gx + (((size_t)b * C + cc) * H + gy) * W + gxx;
        if (gxx < W)
            *reinterpret_cast<float4 *>(o) =
                make_float4(acc[c][0] * rc, acc[c][1] * rc, acc[c][2] * rc, acc[c][3] * rc);
        if (gxx + 4 < W)
            *reinterpret_cast<float4 *>(o + 4) =
                make_float4(acc[c][4] * rc, acc[c][5] * rc, acc[c][6] * rc, acc[c][7] * rc);
    }
}

// mixed launch, 1-D grid: CTAs [0, n2) compute gx2 on 8x32 tiles, the rest gx1 on 16x16 tiles
__global__ void __launch_bounds__(T_THREADS, 2)
corr_bwd_gather_tma_mixed(const __grid_constant__ CUtensorMap mx1, const __grid_constant__ CUtensorMap mg1,
                          const __grid_constant__ CUtensorMap mx2, const __grid_constant__ CUtensorMap mg2,
                          float *__restrict__ gx1, float *__restrict__ gx2, int C, int H, int W, int ngroups, int nz,
                          int t32x, int t32y, int t16x, int t16y, float rc) {
    pdl_wait();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    int id = blockIdx.x;
    const int n2 = t32x * t32y * nz;
    if (id < n2) {
        const int tx = id % t32x, ty = (id / t32x) % t32y, z = id / (t32x * t32y);
        bwd_tile<true, CT>(smem_raw, mx2, mg2, gx2, C, H, W, ngroups, z, tx, ty, rc);
    } else {
        id -= n2;
        const int tx = id % t16x, ty = (id / t16x) % t16y, z = id / (t16x * t16y);
        bwd_tile_sq(smem_raw, mx1, mg1, gx1, C, H, W, ngroups, z, tx, ty, rc);
    }
}

template <int CTT>
cudaError_t launch_tma(const CUtensorMap &mx1, const CUtensorMap &mg1, const CUtensorMap &mx2,
                       const CUtensorMap &mg2, float *gx1, float *gx2, int B, int C, int H, int W,
                       cudaStream_t stream) {
    constexpr int SM = GTile<true, CTT>::kSmem > GTile<false, CTT>::kSmem ? GTile<true, CTT>::kSmem
                                                                          : GTile<false, CTT>::kSmem;
    if (cudaError_t once = ensure_smem((const void *)corr_bwd_gather_tma<CTT>, SM)) return once;
    const int ngroups = (C + NCW * CTT - 1) / (NCW * CTT);
    const int nz = B * ngroups;
    dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, nz * ((gx1 && gx2) ? 2 : 1));
    usf_launch(corr_bwd_gather_tma<CTT>, dim3(grid), dim3(T_THREADS), SM, stream, mx1, mg1, mx2, mg2, gx1, gx2, C, H, W, ngroups, nz,
                                                              1.0f / (float)C);
    return cudaGetLastError();
}

// USF_BWD_MIXED_TILES=0 keeps both gradients on 8x32 tiles (A/B runs, fallback tests)
inline bool mixed_enabled() {
    static const bool on = [] {
        const char *e = getenv("USF_BWD_MIXED_TILES");
        return !(e && e[0] == '0');
    }();
    return on;
}

// CTAs of the 8-channel-per-thread configuration; below ~3/4 of a wave (2 CTAs/SM) the 4-channel one is used
inline bool thin_grid(int B, int C, int H, int W, int ngrad) {
    static const int force = [] { const char *e = getenv("USF_BWD_THIN"); return e ? atoi(e) : -1; }();
    if (force >= 0) return force != 0;
    const long ctas = (long)((W + TW - 1) / TW) * ((H + TH - 1) / TH) * B * ((C + CG - 1) / CG) * ngrad;
    return ctas * 4 < (long)sm_count() * 2 * 3;
}

template <bool GX2, bool VEC>
cudaError_t launch(const float *xin, const float *gout, float *gx, int B, int C, int H, int W,
                   cudaStream_t stream) {
    if (cudaError_t once = ensure_smem((const void *)corr_bwd_gather<GX2, VEC>, SMEM_BYTES)) return once;
    const int ngroups = (C + CG - 1) / CG;
    dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, B * ngroups);
    corr_bwd_gather<GX2, VEC><<<grid, NTHREADS, SMEM_BYTES, stream>>>(xin, gout, gx, C, H, W, ngroups,
                                                                      1.0f / (float)C);
    return cudaGetLastError();
}

}  // namespace

int corr_bwd_direct_f32(const float *, const float *, const float *, float *, float *, int, int, int, int,
                        cudaStream_t);

int corr_bwd_f32(const float *x1, const float *x2, const float *gout, float *gx1, float *gx2, int B, int C,
                 int H, int W, cudaStream_t stream) {
    // (widths TMA can address go to the tiled kernels even when small)
    if ((long)B * H * W <= kDirectMaxPixels && (W % 4) != 0) return corr_bwd_direct_f32(x1, x2, gout, gx1, gx2, B, C, H, W, stream);
    const bool vec = (W % 4 == 0) &&
                     (((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)gout | (uintptr_t)gx1 | (uintptr_t)gx2) % 16 == 0);
    int st = USF_OK;
    if (vec) {
        // TMA path (tensor maps encoded per call, passed as __grid_constant__ parameters)
        CUtensorMap mx1 = {}, mg1 = {}, mx2 = {}, mg2 = {};
        bool ok = true;
        const bool thin = thin_grid(B, C, H, W, (gx1 && gx2) ? 2 : 1);
        if (!thin && gx1 && gx2 && mixed_enabled()) {
            // gx1 on 16x16 tiles when that saves a round of resident CTAs (2 per SM)
            const int ngroups = (C + CG - 1) / CG, nz = B * ngroups;
            const int t32x = (W + TW - 1) / TW, t32y = (H + TH - 1) / TH;
            const int t16x = (W + SQ - 1) / SQ, t16y = (H + SQ - 1) / SQ;
            const long slots = (long)sm_count() * 2, now = 2L * t32x * t32y * nz, mix = (long)(t32x * t32y + t16x * t16y) * nz;
            if ((mix + slots - 1) / slots < (now + slots - 1) / slots &&
                make_nchw_map(&mx1, x2, B, C, H, W, XS2, XR2, CT) && make_nchw_map(&mg1, gout, B, kK, H, W, GS2, SQ, kD) &&
                make_nchw_map(&mx2, x1, B, C, H, W, XS, XR, CT) && make_nchw_map(&mg2, gout, B, kK, H, W, XS, TH, kD)) {
                constexpr int SM = GTile<true, CT>::kSmem > SQ_SMEM ? GTile<true, CT>::kSmem : SQ_SMEM;
                if (cudaError_t once = ensure_smem((const void *)corr_bwd_gather_tma_mixed, SM)) return note_launch(once);
                usf_launch(corr_bwd_gather_tma_mixed, dim3((unsigned)mix), dim3(T_THREADS), SM, stream, 
                    mx1, mg1, mx2, mg2, gx1, gx2, C, H, W, ngroups, nz, t32x, t32y, t16x, t16y, 1.0f / (float)C);
                return note_launch(cudaGetLastError());
            }
        }
        const int ctt = thin ? 4 : CT;
        if (gx1) ok = make_nchw_map(&mx1, x2, B, C, H, W, XS, XR, ctt) && make_nchw_map(&mg1, gout, B, kK, H, W, GS, TH, kD);
        if (ok && gx2) ok = make_nchw_map(&mx2, x1, B, C, H, W, XS, XR, ctt) && make_nchw_map(&mg2, gout, B, kK, H, W, XS, TH, kD);
        if (ok) {
            if (!gx1) { mx1 = mx2; mg1 = mg2; }
            if (!gx2) { mx2 = mx1; mg2 = mg1; }
            return note_launch(thin ? launch_tma<4>(mx1, mg1, mx2, mg2, gx1, gx2, B, C, H, W, stream)
                                    : launch_tma<8>(mx1, mg1, mx2, mg2, gx1, gx2, B, C, H, W, stream));
        }
        // encoding failed (driver entry point missing / TMA disabled): fall through to the cp.async kernels
    }
    if (gx1) {
        cudaError_t e = vec ? launch<false, true>(x2, gout, gx1, B, C, H, W, stream)
                            : launch<false, false>(x2, gout, gx1, B, C, H, W, stream);
        st = note_launch(e);
        if (st != USF_OK) return st;
    }
    if (gx2) {
        cudaError_t e = vec ? launch<true, true>(x1, gout, gx2, B, C, H, W, stream)
                            : launch<true, false>(x1, gout, gx2, B, C, H, W, stream);
        st = note_launch(e);
    }
    return st;
}

}  // namespace usf
