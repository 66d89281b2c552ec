// corr_small.cu -- cost volume fwd/bwd for pyramid levels whose WHOLE image is at most 256 pixels.
//
// The coarsest PWC-Lite levels (KITTI 4x13 and 8x26, Sintel 7x16; models/pwclite.py:295 levels 0-1)
// have odd widths (rows are not 16-byte multiples: no row-wise TMA boxes, no row-wise vectors) and a
// few hundred pixels per image, so the 8x32 tiling of corr_fwd.cu / corr_bwd.cu leaves most of the GPU
// idle.  But a channel's whole image is ONE contiguous run of H*W floats, so these kernels keep images
// in shared memory PACKED, exactly as they lie in global memory, and index them with p = y*W + x:
//   * fill: fp32 images with H*W % 4 == 0 come in through the TMA unit -- fwd: one 2-D box
//     [channels][row block] per operand over the [B*C][H*W] view of the tensor; bwd: two 1-D bulk copies
//     (all 81 channels of gO[b], the CTA's channel slice of x) -- completion on an mbarrier, no per-thread
//     index math, no register staging.  bf16 / other shapes use batched LDG + STS;
//   * a displaced position whose COLUMN falls outside the image wraps into a neighbouring row of the
//     packed layout.  Nothing is predicated in the inner loops: the forward computes those (pixel, dx)
//     sums on wrapped data and discards them at the store (they are exact zeros in the reference), the
//     backward zeroes the gO factor of such terms once per dy (18 selects per 144 FFMA).  Displaced ROWS
//     outside the image are skipped.  (The wrapped operand is real neighbouring data, so a non-finite
//     feature value next to the image border can leak into a border gradient as 0 * inf; finite inputs
//     -- the only case the reference's own training loop survives -- are exact.)
//   * even widths run two pixels per thread (PX = 2): the x2 window of a pixel pair is 5 aligned
//     LDS.64 at constant offsets, 18 FFMA per 6 loads;
//   * fwd: CTA = (image, dy, row block); channel groups of the CTA reduce through shared memory in a
//     fixed order (deterministic).  bwd: CTA = (image, 8 or 16 channels, gradient) holds the image's whole
//     81-channel gO (the same gather out[c,p] = sum_e Gsel(e,p) * X[c,p+e] as corr_bwd.cu, see the
//     derivation at the top of that file).
// Formulas: models/correlation_native.py:13-23.  Storage T is fp32 or bf16, arithmetic fp32.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "tma.cuh"

namespace usf {

// images of at most this many pixels take the kernels in this file (USF_SMALL_MAX_HW=0 disables them)
long kSmallMaxHW = [] {
    const char *e = getenv("USF_SMALL_MAX_HW");
    const long v = e ? atol(e) : 256L;
    return v > 256L ? 256L : v;
}();

namespace {

constexpr int kPad = 32;  // floats (128 bytes) of slack before and after the shared-memory arrays
constexpr int kMaxThreads = 512;

// n contiguous elements -> shared, U loads in flight per thread before the first store
template <typename T>
__device__ __forceinline__ void load_flat(float *__restrict__ dst, const T *__restrict__ src, int n, int tid,
                                          int nthr) {
    constexpr int U = 8;
    for (int i0 = tid; i0 < n; i0 += U * nthr) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i0 + u * nthr < n) v[u] = ldf(src + i0 + u * nthr);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i0 + u * nthr < n) dst[i0 + u * nthr] = v[u];
    }
}

// nslabs runs of n contiguous elements (global stride gstride) -> shared slabs of stride S floats.
// Flat (slab, element) index per thread, advanced incrementally: two divisions per call, none per element.
template <typename T>
__device__ __forceinline__ void load_slabs(float *__restrict__ dst, int S, const T *__restrict__ src, size_t gstride,
                                           int nslabs, int n, int tid, int nthr) {
    constexpr int U = 8;
    const int total = nslabs * n;
    const int ds = nthr / n, dj = nthr - ds * n;
    int sl = tid / n, j = tid - sl * n;
    for (int i = tid; i < total; i += U * nthr) {
        float v[U];
        int off[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            off[u] = -1;
            if (i + u * nthr < total) {
                off[u] = sl * S + j;
                v[u] = ldf(src + (size_t)sl * gstride + j);
            }
            j += dj;
            sl += ds;
            if (j >= n) {
                j -= n;
                ++sl;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (off[u] >= 0) dst[off[u]] = v[u];
    }
}

template <int PX, typename T>
__device__ __forceinline__ void store_px(T *p, const float *v) {
    if (PX == 2 && std::is_same<T, float>::value) {
        *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
        for (int i = 0; i < PX; ++i) stf(p + i, v[i]);
    }
}

// bit w set <=> window column x - 4 + w lies inside the image
__device__ __forceinline__ unsigned window_mask(int x, int W, int n) {
    unsigned m = 0;
    for (int w = 0; w < n; ++w) {
        const int col = x - kMaxDisp + w;
        if (col >= 0 && col < W) m |= 1u << w;
    }
    return m;
}

// ---------------------------------------------------------------------------------------------
// forward.  grid = (nblk * 9, B), block = NG << lgP threads: thread (g, pp) owns PX pixels x 9 dx of the
// CTA's dy for the channels c = g, g + NG, ...
// shared (floats): [kPad] x1s[CK][S] .. x2s[CK][S] [kPad]; slab c of x1s holds elements [lo1, lo1+S) of
// channel image c0+c, slab c of x2s elements [lo2, lo2+S) -- lo = first needed element, rounded down to a
// 16-byte boundary on the TMA path (out-of-image parts of a box arrive as zeros).  The arrays are reused
// for the cross-group reduction at the end.
// ---------------------------------------------------------------------------------------------
template <int PX, bool LEAKY, typename T>
__global__ void __launch_bounds__(kMaxThreads)
corr_fwd_small(const __grid_constant__ CUtensorMap m1, const __grid_constant__ CUtensorMap m2,
               const T *__restrict__ x1, const T *__restrict__ x2, T *__restrict__ out, int C, int H, int W, int RB,
               int lgP, int NG, int CK, int S, int tma, float rc, float slope, size_t obs) {
    pdl_wait();
    extern __shared__ __align__(1024) float sm[];
    __shared__ uint64_t fill_bar;
    const int tid = threadIdx.x, nthr = blockDim.x, P = 1 << lgP;
    const int dyi = blockIdx.x % kD, blk = blockIdx.x / kD, b = blockIdx.y;
    const int dy = dyi - kMaxDisp, y0 = blk * RB;
    const int rows = min(RB, H - y0), HW = H * W;
    const int r_lo = max(0, -(y0 + dy)), r_hi = min(rows, H - (y0 + dy));  // local rows whose x2 row exists
    const int g = tid >> lgP, pp = tid & (P - 1), p0 = pp * PX;
    const bool active = p0 < rows * W;
    const int r = p0 / W, x = p0 - r * W;
    const bool rowok = active && r >= r_lo && r < r_hi;
    T *optr = out + (size_t)b * obs + (((size_t)dyi * kD) * H + y0 + r) * W + x;

    if (r_hi <= r_lo) {  // every displaced row is outside the image: the CTA's outputs are zeros
        if (active) {
            const float z[2] = {0.f, 0.f};
            for (int dx = g; dx < kD; dx += NG) store_px<PX, T>(optr + (size_t)dx * HW, z);
        }
        return;
    }

    int lo1, lo2, n1, n2;
    if (tma) {
        lo1 = (y0 * W) & ~3;
        lo2 = ((y0 + dy) * W) & ~3;  // may be negative: rounds toward -inf, the box is zero-filled there
        n1 = n2 = S;
    } else {
        lo1 = y0 * W;
        n1 = rows * W;
        lo2 = (y0 + dy + r_lo) * W;
        n2 = (r_hi - r_lo) * W;
    }
    const int off1 = y0 * W - lo1, off2 = (y0 + dy) * W - lo2;
    const int x2off = (CK * S + 31) & ~31;  // keeps x2s 128-byte aligned
    float *x1s = sm + kPad, *x2s = x1s + x2off;

    float acc[kD][PX];
#pragma unroll
    for (int dx = 0; dx < kD; ++dx)
#pragma unroll
        for (int i = 0; i < PX; ++i) acc[dx][i] = 0.f;

    if (tma) {
        if (tid == 0) {
            mbar_init(&fill_bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    uint32_t phase = 0;
    for (int c0 = 0; c0 < C; c0 += CK) {
        const int ck = min(CK, C - c0);
        if (c0) {  // the slabs are about to be overwritten (by TMA on the fast path): reads must have completed
            __threadfence_block();
            __syncthreads();
        }
        if (tma) {
            if (tid == 0) {
                mbar_expect_tx(&fill_bar, 2u * CK * S * 4u);
                tma_load_2d(x1s, &m1, &fill_bar, lo1, b * C + c0);
                tma_load_2d(x2s, &m2, &fill_bar, lo2, b * C + c0);
            }
            mbar_wait(&fill_bar, phase);
            phase ^= 1;
        } else {
            load_slabs(x1s, S, x1 + ((size_t)b * C + c0) * HW + lo1, (size_t)HW, ck, n1, tid, nthr);
            load_slabs(x2s, S, x2 + ((size_t)b * C + c0) * HW + lo2, (size_t)HW, ck, n2, tid, nthr);
            __syncthreads();
        }
        if (rowok) {
            const float *p1 = x1s + p0 + off1, *p2 = x2s + p0 + off2 - kMaxDisp;
#pragma unroll 2
            for (int c = g; c < ck; c += NG) {
                if (PX == 1) {
                    const float a = p1[c * S];
#pragma unroll
                    for (int dx = 0; dx < kD; ++dx) acc[dx][0] = fmaf(a, p2[c * S + dx], acc[dx][0]);
                } else {
                    const float2 a = *reinterpret_cast<const float2 *>(p1 + c * S);
                    float w[10];
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        const float2 v = *reinterpret_cast<const float2 *>(p2 + c * S + 2 * j);
                        w[2 * j] = v.x;
                        w[2 * j + 1] = v.y;
                    }
#pragma unroll
                    for (int dx = 0; dx < kD; ++dx) {
                        acc[dx][0] = fmaf(a.x, w[dx], acc[dx][0]);
                        acc[dx][PX - 1] = fmaf(a.y, w[dx + 1], acc[dx][PX - 1]);
                    }
                }
            }
        }
    }

    // fixed-order reduction over the NG channel groups: red[g][dx*PX + i][pp]
    __syncthreads();
    float *red = x1s;
#pragma unroll
    for (int dx = 0; dx < kD; ++dx)
#pragma unroll
        for (int i = 0; i < PX; ++i) red[((g * kD + dx) * PX + i) * P + pp] = acc[dx][i];
    __syncthreads();
    if (!active) return;
    const unsigned vmask = rowok ? window_mask(x, W, kD + PX - 1) : 0u;
    for (int dx = g; dx < kD; dx += NG) {
        float v[PX];
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            float s = red[(dx * PX + i) * P + pp];
            for (int gg = 1; gg < NG; ++gg) s += red[((gg * kD + dx) * PX + i) * P + pp];
            s *= rc;
            if (LEAKY) s = s < 0.f ? s * slope : s;
            v[i] = ((vmask >> (dx + i)) & 1u) ? s : 0.f;  // wrapped columns: exact zeros in the reference
        }
        store_px<PX, T>(optr + (size_t)dx * HW, v);
    }
}

// ---------------------------------------------------------------------------------------------
// backward.  grid = (ceil(C / CS), B, gradients), block = NGC << lgP threads; thread (g, pp) owns PX pixels
// x 8 channels.  shared (floats): [kPad] Gs[81][HW] Xs[CS][HW] [kPad], both packed like global memory.
// ---------------------------------------------------------------------------------------------
constexpr int kCT = 8;  // channels per thread

template <int PX, bool GX2>
__device__ __forceinline__ void bwd_small_body(float (&acc)[kCT][PX], const float *__restrict__ Gs,
                                               const float *__restrict__ Xs, int HW, int H, int W, int p0) {
    const int y = p0 / W, x = p0 - y * W;
    const unsigned vmask = window_mask(x, W, kD + PX - 1);
#pragma unroll 1
    for (int dyi = 0; dyi < kD; ++dyi) {
        const int ey = dyi - kMaxDisp, yy = y + ey;
        if (yy < 0 || yy >= H) continue;
        const int base = p0 + ey * W - kMaxDisp;  // packed index of window element 0 (column x-4)
        float g[kD][PX];
#pragma unroll
        for (int dxi = 0; dxi < kD; ++dxi) {
            const int k = dyi * kD + dxi;
            float v[2];
            if (!GX2) {
                const float *gk = Gs + k * HW + p0;
                if (PX == 2) {
                    const float2 t = *reinterpret_cast<const float2 *>(gk);
                    v[0] = t.x;
                    v[1] = t.y;
                } else {
                    v[0] = v[1] = gk[0];
                }
            } else {
                const float *gk = Gs + (kK - 1 - k) * HW + base + dxi;
                if (PX == 2 && (dxi & 1) == 0) {
                    const float2 t = *reinterpret_cast<const float2 *>(gk);
                    v[0] = t.x;
                    v[1] = t.y;
                } else {
                    v[0] = gk[0];
                    v[1] = gk[PX - 1];
                }
            }
#pragma unroll
            for (int i = 0; i < PX; ++i) g[dxi][i] = ((vmask >> (dxi + i)) & 1u) ? v[i] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < kCT; ++j) {
            const float *xw = Xs + j * HW + base;
            if (PX == 1) {
#pragma unroll
                for (int dxi = 0; dxi < kD; ++dxi) acc[j][0] = fmaf(g[dxi][0], xw[dxi], acc[j][0]);
            } else {
                float w[10];
#pragma unroll
                for (int t = 0; t < 5; ++t) {
                    const float2 u = *reinterpret_cast<const float2 *>(xw + 2 * t);
                    w[2 * t] = u.x;
                    w[2 * t + 1] = u.y;
                }
#pragma unroll
                for (int dxi = 0; dxi < kD; ++dxi) {
                    acc[j][0] = fmaf(g[dxi][0], w[dxi], acc[j][0]);
                    acc[j][PX - 1] = fmaf(g[dxi][PX - 1], w[dxi + 1], acc[j][PX - 1]);
                }
            }
        }
    }
}

template <int PX, typename T>
__global__ void __launch_bounds__(kMaxThreads)
corr_bwd_small(const T *__restrict__ x1, const T *__restrict__ x2, const T *__restrict__ gout, T *__restrict__ gx1,
               T *__restrict__ gx2, int C, int H, int W, int lgP, int NGC, int bulk, float rc) {
    pdl_wait();
    extern __shared__ __align__(1024) float sm[];
    __shared__ uint64_t fill_bar;
    const int tid = threadIdx.x, nthr = blockDim.x, P = 1 << lgP;
    const int HW = H * W, CS = NGC * kCT;
    const int c0 = blockIdx.x * CS, b = blockIdx.y, ck = min(CS, C - c0);
    const bool second = blockIdx.z == 1 || !gx1;  // this CTA computes gx2 (gathers over x1), else gx1 (over x2)
    float *Gs = sm + kPad, *Xs = Gs + kK * HW;

    const T *sg = gout + (size_t)b * kK * HW, *sx = (second ? x1 : x2) + ((size_t)b * C + c0) * HW;
    if (bulk) {
        if (tid == 0) {
            mbar_init(&fill_bar, 1);
            mbar_fence_init();
            mbar_expect_tx(&fill_bar, (uint32_t)(kK + ck) * HW * 4);
            bulk_load_1d(Gs, sg, (uint32_t)kK * HW * 4, &fill_bar);
            bulk_load_1d(Xs, sx, (uint32_t)ck * HW * 4, &fill_bar);
        }
    } else {
        load_flat(Gs, sg, kK * HW, tid, nthr);
        load_flat(Xs, sx, ck * HW, tid, nthr);
    }
    // wrapped reads reach a few floats past either end of a slab: keep everything they can touch finite
    for (int i = tid; i < kPad; i += nthr) sm[i] = Xs[CS * HW + i] = 0.f;
    for (int i = ck * HW + tid; i < CS * HW; i += nthr) Xs[i] = 0.f;
    __syncthreads();  // also publishes the mbarrier init
    if (bulk) mbar_wait(&fill_bar, 0);

    const int g = tid >> lgP, pp = tid & (P - 1), p0 = pp * PX;
    if (p0 >= HW || g * kCT >= ck) return;
    float acc[kCT][PX];
#pragma unroll
    for (int j = 0; j < kCT; ++j)
#pragma unroll
        for (int i = 0; i < PX; ++i) acc[j][i] = 0.f;
    if (second)
        bwd_small_body<PX, true>(acc, Gs, Xs + g * kCT * HW, HW, H, W, p0);
    else
        bwd_small_body<PX, false>(acc, Gs, Xs + g * kCT * HW, HW, H, W, p0);

    T *o = (second ? gx2 : gx1) + ((size_t)b * C + c0 + g * kCT) * HW + p0;
#pragma unroll
    for (int j = 0; j < kCT; ++j) {
        if (g * kCT + j >= ck) break;
        float v[PX];
#pragma unroll
        for (int i = 0; i < PX; ++i) v[i] = acc[j][i] * rc;
        store_px<PX, T>(o + (size_t)j * HW, v);
    }
}

// threads per channel group: next power of two >= max(32, n)
inline int log2_threads(int n) {
    int lg = 5;
    while ((1 << lg) < n) ++lg;
    return lg;
}

// the TMA fills need fp32 storage, 16-byte aligned channel images and the TMA unit enabled
template <typename T>
bool tma_fill_ok(int HW, const void *a, const void *b, const void *c) {
    const uintptr_t bits = (uintptr_t)a | (uintptr_t)b | (uintptr_t)c;
    return std::is_same<T, float>::value && (HW % 4) == 0 && (bits & 15) == 0 && tma_enabled();
}

// [rows][cols] fp32 view, box [bh][bw]
inline bool make_rows_map(CUtensorMap *map, const void *base, long rows, int cols, int bw, int bh) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || bw > 256 || bh > 256) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int PX, typename T>
cudaError_t launch_fwd(const T *x1, const T *x2, T *out, int B, int C, int H, int W, unsigned flags, float slope,
                       cudaStream_t stream) {
    // row blocks: about two CTAs per SM when the batch alone does not provide them (smaller fills, fill of one
    // CTA overlaps the math of the other), but never so few rows that most lanes of a warp idle
    static const int target = [] { const char *e = getenv("USF_SMALL_FWD_CTAS"); return e ? atoi(e) : 296; }();
    int nblk = (target + B * kD - 1) / (B * kD);
    nblk = nblk < 1 ? 1 : (nblk > H ? H : nblk);
    int RB = (H + nblk - 1) / nblk;
    const int rb_min = (24 * PX + W - 1) / W;
    if (RB < rb_min) RB = rb_min < H ? rb_min : H;
    nblk = (H + RB - 1) / RB;
    RB = (H + nblk - 1) / nblk;
    const int lgP = log2_threads((RB * W + PX - 1) / PX), P = 1 << lgP;
    int NG = kMaxThreads / P;
    NG = NG > 8 ? 8 : (NG < 1 ? 1 : NG);
    while (NG > 1 && NG * 4 > C) --NG;
    int tma = tma_fill_ok<T>(H * W, x1, x2, x1);
    int S = tma ? ((RB * W + 3 + 3) & ~3) : RB * W;  // + up to 3 floats of 16-byte rounding at the front
    const int budget = 110 * 1024;
    int CK = budget / (2 * S * 4);
    CK = CK > C ? C : CK;
    CK = CK > 256 ? 256 : CK;
    CUtensorMap m1 = {}, m2 = {};
    if (tma)
        tma = S <= 256 && make_rows_map(&m1, x1, (long)B * C, H * W, S, CK) &&
              make_rows_map(&m2, x2, (long)B * C, H * W, S, CK);
    if (!tma) S = RB * W;
    const size_t x2off = ((size_t)CK * S + 31) & ~(size_t)31;
    size_t fl = 2 * x2off;
    const size_t red = (size_t)NG * kD * PX * P;
    if (red > fl) fl = red;
    const size_t smem = (fl + 2 * kPad) * 4;
    const float rc = 1.0f / (float)C;
    dim3 grid(nblk * kD, B);
    cudaError_t e;
    if (flags & USF_FLAG_LEAKY_RELU) {
        if ((e = ensure_smem((const void *)corr_fwd_small<PX, true, T>, (int)smem)) != cudaSuccess) return e;
        usf_launch((corr_fwd_small<PX, true, T>), dim3(grid), dim3(P * NG), smem, stream, m1, m2, x1, x2, out, C, H, W, RB, lgP, NG, CK, S,
                                                                    tma, rc, slope, corr_out_bstride(H, W));
    } else {
        if ((e = ensure_smem((const void *)corr_fwd_small<PX, false, T>, (int)smem)) != cudaSuccess) return e;
        usf_launch((corr_fwd_small<PX, false, T>), dim3(grid), dim3(P * NG), smem, stream, m1, m2, x1, x2, out, C, H, W, RB, lgP, NG, CK, S,
                                                                     tma, rc, slope, corr_out_bstride(H, W));
    }
    return cudaGetLastError();
}

template <int PX, typename T>
cudaError_t launch_bwd(const T *x1, const T *x2, const T *gout, T *gx1, T *gx2, int B, int C, int H, int W,
                       cudaStream_t stream) {
    const int HW = H * W;
    const int lgP = log2_threads((HW + PX - 1) / PX), P = 1 << lgP;
    const int ngrad = (gx1 && gx2) ? 2 : 1;
    // 16 channels per CTA halve the copies of gO; 8 per CTA when that is what it takes to cover the SMs
    int NGC = kMaxThreads / P;
    NGC = NGC > 2 ? 2 : (NGC < 1 ? 1 : NGC);
    if (C <= kCT || (long)B * ngrad * ((C + 2 * kCT - 1) / (2 * kCT)) < sm_count()) NGC = 1;
    const int CS = NGC * kCT;
    const size_t smem = (((size_t)kK + CS) * HW + 2 * kPad) * 4;
    cudaError_t e = ensure_smem((const void *)corr_bwd_small<PX, T>, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((C + CS - 1) / CS, B, ngrad);
    usf_launch((corr_bwd_small<PX, T>), dim3(grid), dim3(NGC * P), smem, stream, x1, x2, gout, gx1, gx2, C, H, W, lgP, NGC,
                                                           tma_fill_ok<T>(HW, x1, x2, gout), 1.0f / (float)C);
    return cudaGetLastError();
}

template <typename T>
bool pairs_ok(int W, const void *a, const void *b, const void *c) {
    // two pixels per thread need even widths; fp32 pair stores need 8-byte aligned bases
    const uintptr_t bits = (uintptr_t)a | (uintptr_t)b | (uintptr_t)c;
    return (W % 2 == 0) && (bits & 7) == 0;
}

}  // namespace

bool corr_small_applies(int H, int W) { return (long)H * W <= kSmallMaxHW; }

template <typename T>
static int fwd_small(const T *x1, const T *x2, T *out, int B, int C, int H, int W, unsigned flags, float slope,
                     cudaStream_t stream) {
    if (pairs_ok<T>(W, x1, x2, out) && corr_out_bstride(H, W) % 2 == 0) return note_launch(launch_fwd<2, T>(x1, x2, out, B, C, H, W, flags, slope, stream));
    return note_launch(launch_fwd<1, T>(x1, x2, out, B, C, H, W, flags, slope, stream));
}

template <typename T>
static int bwd_small(const T *x1, const T *x2, const T *gout, T *gx1, T *gx2, int B, int C, int H, int W,
                     cudaStream_t stream) {
    if (pairs_ok<T>(W, gout, gx1, gx2)) return note_launch(launch_bwd<2, T>(x1, x2, gout, gx1, gx2, B, C, H, W, stream));
    return note_launch(launch_bwd<1, T>(x1, x2, gout, gx1, gx2, B, C, H, W, stream));
}

int corr_fwd_small_f32(const float *x1, const float *x2, float *out, int B, int C, int H, int W, unsigned flags,
                       float slope, cudaStream_t stream) {
    return fwd_small<float>(x1, x2, out, B, C, H, W, flags, slope, stream);
}
int corr_bwd_small_f32(const float *x1, const float *x2, const float *gout, float *gx1, float *gx2, int B, int C,
                       int H, int W, cudaStream_t stream) {
    return bwd_small<float>(x1, x2, gout, gx1, gx2, B, C, H, W, stream);
}
int corr_fwd_small_bf16(const void *x1, const void *x2, void *out, int B, int C, int H, int W, unsigned flags,
                        float slope, cudaStream_t stream) {
    return fwd_small<bf16>((const bf16 *)x1, (const bf16 *)x2, (bf16 *)out, B, C, H, W, flags, slope, stream);
}
int corr_bwd_small_bf16(const void *x1, const void *x2, const void *gout, void *gx1, void *gx2, int B, int C, int H,
                        int W, cudaStream_t stream) {
    return bwd_small<bf16>((const bf16 *)x1, (const bf16 *)x2, (const bf16 *)gout, (bf16 *)gx1, (bf16 *)gx2, B, C, H,
                           W, stream);
}

}  // namespace usf
