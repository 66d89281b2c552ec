// corr_fwd.cu -- cost-volume forward for sm_100a (fp32 storage, fp32 FFMA).
//
// Replaces: correlation_cuda.forward (models/correlation_package/correlation_cuda.cc:10-86
// -> channels_first x2 + correlation_forward, correlation_cuda_kernel.cu:15-114), same result
// as models/correlation_native.py:13-23:
//     out[b,(dy+4)*9+(dx+4),y,x] = (1/C) sum_c x1[b,c,y,x] * x2[b,c,y+dy,x+dx]
//
// Design (one launch, no NHWC scratch, no fill, no atomics):
//   * a CTA owns an 8x32 pixel tile; channels stream through shared memory in chunks of 8
//     with a 3-stage cp.async (LDGSTS) ring; the x2 tile carries its 4-pixel halo, and image
//     borders / ragged channel tails are zero-filled by cp.async src-size 0 (== F.pad zeros);
//   * a thread owns 8 consecutive pixels of one row and the 9 dx of ONE dy: 72 fp32
//     accumulators; per channel it issues 2 + 4 LDS.128 (x1 strip, 16-wide x2 window) for
//     72 FFMA.  Row strides of 9 and 11 16-byte chunks make every LDS.128 phase conflict-free;
//   * when W % 4 == 0 the stages are filled by TMA (cp.async.bulk.tensor 4-D boxes, hardware
//     zero-fill for halo / ragged tails) issued by one producer thread and handed to the consumer
//     warps through full/empty mbarriers: no address arithmetic and no __syncthreads in the loop;
//     the cp.async kernels below remain as the path for odd widths (W = 13, 26, ...);
//   * kernel A ("tile9"): 9 warps = 9 dy, every warp walks all channels (big levels);
//   * kernel B ("splitc"): CTA = (tile, dy), 8 warps take every 8th channel and their partial
//     tiles are summed through shared memory in a fixed order (small levels: few pixels, many
//     channels -> 9x more CTAs, no y-halo).  Deterministic either way.
#include "common.cuh"
#include "tma.cuh"

namespace usf {
namespace {

constexpr int TH = 8;    // tile rows
constexpr int TW = 32;   // tile cols
constexpr int CK = 8;    // channels per pipeline stage
constexpr int NST = 3;   // pipeline stages
constexpr int X1S = 36;  // x1 tile row stride in floats (9 chunks: odd -> conflict-free)
constexpr int X2W = TW + 2 * kMaxDisp;  // 40 halo columns
constexpr int X2S = 44;  // x2 tile row stride in floats (11 chunks)

// ---------------------------------------------------------------------------
// Stage loader.  ROWS2 = number of x2 rows staged per channel, first row = gy2_0.
// VEC: W % 4 == 0 and 16-byte aligned bases -> every 4-float chunk is entirely
// inside or outside the image, so one 16-byte cp.async with zero-fill does it.
// ---------------------------------------------------------------------------
template <int ROWS2, bool VEC, int NT>
__device__ __forceinline__ void load_stage(float *__restrict__ st, const float *__restrict__ x1b,
                                           const float *__restrict__ x2b, int c0, int C, int H, int W,
                                           int y0, int x0, int gy2_0, int tid) {
    constexpr int X1_CHUNKS = CK * TH * (TW / 4);  // 512
    float *st2 = st + CK * TH * X1S;
    for (int i = tid; i < X1_CHUNKS; i += NT) {
        const int c = i >> 6, r = (i >> 3) & 7, ch = i & 7;
        const int cc = c0 + c, gy = y0 + r, gx = x0 + 4 * ch;
        float *dst = st + (c * TH + r) * X1S + 4 * ch;
        const bool rowok = (cc < C) && (gy < H);
        const float *src = x1b + ((size_t)cc * H + gy) * W + gx;
        if (VEC) {
            const bool ok = rowok && (gx < W);
            cp_async16(dst, ok ? src : x1b, ok ? 16 : 0);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool ok = rowok && (gx + e < W);
                cp_async4(dst + e, ok ? src + e : x1b, ok ? 4 : 0);
            }
        }
    }
    constexpr int X2_CHUNKS = CK * ROWS2 * (X2W / 4);
    for (int i = tid; i < X2_CHUNKS; i += NT) {
        const int c = i / (ROWS2 * 10);
        const int rem = i - c * (ROWS2 * 10);
        const int r = rem / 10, ch = rem - r * 10;
        const int cc = c0 + c, gy = gy2_0 + r, gx = x0 - kMaxDisp + 4 * ch;
        float *dst = st2 + (c * ROWS2 + r) * X2S + 4 * ch;
        const bool rowok = (cc < C) && (gy >= 0) && (gy < H);
        const float *src = x2b + ((size_t)cc * H + gy) * W + gx;
        if (VEC) {
            const bool ok = rowok && (gx >= 0) && (gx < W);
            cp_async16(dst, ok ? src : x2b, ok ? 16 : 0);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool ok = rowok && (gx + e >= 0) && (gx + e < W);
                cp_async4(dst + e, ok ? src + e : x2b, ok ? 4 : 0);
            }
        }
    }
}

// 72 FFMA for one channel: acc[dx][i] += x1[i] * x2win[dx + i]
__device__ __forceinline__ void fma_channel(float (&acc)[kD][8], const float *__restrict__ p1,
                                            const float *__restrict__ p2) {
    const float4 a0 = lds128(p1), a1 = lds128(p1 + 4);
    const float4 w0 = lds128(p2), w1 = lds128(p2 + 4), w2 = lds128(p2 + 8), w3 = lds128(p2 + 12);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float w[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w,
                         w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
    for (int dx = 0; dx < kD; ++dx)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[dx][i] = fmaf(a[i], w[dx + i], acc[dx][i]);
}

// Operands of one channel for one thread: x1 strip (8) + x2 window (16), as six 16-byte LDS.
struct ChanOps {
    float4 a0, a1, w0, w1, w2, w3;
};
__device__ __forceinline__ void load_ops(ChanOps &o, const float *__restrict__ p1, const float *__restrict__ p2) {
    o.a0 = lds128(p1); o.a1 = lds128(p1 + 4);
    o.w0 = lds128(p2); o.w1 = lds128(p2 + 4); o.w2 = lds128(p2 + 8); o.w3 = lds128(p2 + 12);
}
__device__ __forceinline__ void fma_ops(float (&acc)[kD][8], const ChanOps &o) {
    const float a[8] = {o.a0.x, o.a0.y, o.a0.z, o.a0.w, o.a1.x, o.a1.y, o.a1.z, o.a1.w};
    const float w[16] = {o.w0.x, o.w0.y, o.w0.z, o.w0.w, o.w1.x, o.w1.y, o.w1.z, o.w1.w,
                         o.w2.x, o.w2.y, o.w2.z, o.w2.w, o.w3.x, o.w3.y, o.w3.z, o.w3.w};
#pragma unroll
    for (int dx = 0; dx < kD; ++dx)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[dx][i] = fmaf(a[i], w[dx + i], acc[dx][i]);
}

__device__ __forceinline__ float finish(float v, float rc, bool leaky, float slope) {
    v *= rc;
    return (leaky && v < 0.f) ? v * slope : v;
}

// ---------------------------------------------------------------------------
// Kernel A: 9 warps, warp == dy.  grid = (tiles_x, tiles_y, B)
// ---------------------------------------------------------------------------
constexpr int A_THREADS = 32 * kD;                       // 288
constexpr int A_ROWS2 = TH + 2 * kMaxDisp;               // 16
constexpr int A_STAGE = CK * TH * X1S + CK * A_ROWS2 * X2S;  // 7936 floats
constexpr int A_SMEM = NST * A_STAGE * 4;                // 95232 B

template <bool VEC>
__global__ void __maxnreg__(112)
corr_fwd_tile9(const float *__restrict__ x1, const float *__restrict__ x2, float *__restrict__ out, int C,
               int H, int W, float rc, int leaky, float slope, size_t obs) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = lane >> 2, q = lane & 3;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
    const float *x1b = x1 + (size_t)b * C * H * W;
    const float *x2b = x2 + (size_t)b * C * H * W;
    const int nchunk = (C + CK - 1) / CK;

    float acc[kD][8];
#pragma unroll
    for (int dx = 0; dx < kD; ++dx)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[dx][i] = 0.f;

#pragma unroll
    for (int s = 0; s < NST - 1; ++s) {
        if (s < nchunk)
            load_stage<A_ROWS2, VEC, A_THREADS>(smem + s * A_STAGE, x1b, x2b, s * CK, C, H, W, y0, x0,
                                                y0 - kMaxDisp, tid);
        cp_async_commit();
    }
    for (int k = 0; k < nchunk; ++k) {
        cp_async_wait<NST - 2>();
        __syncthreads();
        const int kn = k + NST - 1;
        if (kn < nchunk)
            load_stage<A_ROWS2, VEC, A_THREADS>(smem + (kn % NST) * A_STAGE, x1b, x2b, kn * CK, C, H, W,
                                                y0, x0, y0 - kMaxDisp, tid);
        cp_async_commit();
        const float *st = smem + (k % NST) * A_STAGE;
        const float *p1 = st + r * X1S + 8 * q;
        const float *p2 = st + CK * TH * X1S + (r + warp) * X2S + 8 * q;
#pragma unroll 2
        for (int c = 0; c < CK; ++c) fma_channel(acc, p1 + c * TH * X1S, p2 + c * A_ROWS2 * X2S);
    }

    const int gy = y0 + r, gx = x0 + 8 * q;
    if (gy >= H) return;
    float *o = out + (size_t)b * obs + (((size_t)warp * kD) * H + gy) * W + gx;
    const size_t kstride = (size_t)H * W;
#pragma unroll
    for (int dx = 0; dx < kD; ++dx, o += kstride) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = finish(acc[dx][i], rc, leaky, slope);
        if (VEC) {
            if (gx < W) *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
            if (gx + 4 < W) *reinterpret_cast<float4 *>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (gx + i < W) o[i] = v[i];
        }
    }
}

// ---------------------------------------------------------------------------
// Kernel B: CTA = (tile, b, dy); 8 warps split the channels; smem tree-less ordered sum.
// grid = (tiles_x, tiles_y, B * 9)
// ---------------------------------------------------------------------------
constexpr int B_WARPS = 8;
constexpr int B_THREADS = 32 * B_WARPS;                  // 256
constexpr int B_STAGE = CK * TH * X1S + CK * TH * X2S;   // 5120 floats
constexpr int B_RED = B_WARPS * kD * TH * X1S;           // 20736 floats (82944 B)
constexpr int B_SMEM = (B_RED > NST * B_STAGE ? B_RED : NST * B_STAGE) * 4;

template <bool VEC>
__global__ void __launch_bounds__(B_THREADS, 2)
corr_fwd_splitc(const float *__restrict__ x1, const float *__restrict__ x2, float *__restrict__ out, int C,
                int H, int W, float rc, int leaky, float slope, size_t obs) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = lane >> 2, q = lane & 3;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int b = blockIdx.z / kD, dyi = blockIdx.z - b * kD;
    const float *x1b = x1 + (size_t)b * C * H * W;
    const float *x2b = x2 + (size_t)b * C * H * W;
    const int nchunk = (C + CK - 1) / CK;
    const int gy2_0 = y0 + dyi - kMaxDisp;

    float acc[kD][8];
#pragma unroll
    for (int dx = 0; dx < kD; ++dx)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[dx][i] = 0.f;

#pragma unroll
    for (int s = 0; s < NST - 1; ++s) {
        if (s < nchunk)
            load_stage<TH, VEC, B_THREADS>(smem + s * B_STAGE, x1b, x2b, s * CK, C, H, W, y0, x0, gy2_0, tid);
        cp_async_commit();
    }
    for (int k = 0; k < nchunk; ++k) {
        cp_async_wait<NST - 2>();
        __syncthreads();
        const int kn = k + NST - 1;
        if (kn < nchunk)
            load_stage<TH, VEC, B_THREADS>(smem + (kn % NST) * B_STAGE, x1b, x2b, kn * CK, C, H, W, y0, x0,
                                           gy2_0, tid);
        cp_async_commit();
        const float *st = smem + (k % NST) * B_STAGE;
        // warp w takes channel w of this chunk (CK == B_WARPS)
        fma_channel(acc, st + (warp * TH + r) * X1S + 8 * q,
                    st + CK * TH * X1S + (warp * TH + r) * X2S + 8 * q);
    }
    cp_async_wait<0>();
    __syncthreads();  // everyone is done with the stage buffers: reuse them for the reduction

    float *red = smem + (size_t)warp * (kD * TH * X1S);
#pragma unroll
    for (int dx = 0; dx < kD; ++dx) {
        float *p = red + (dx * TH + r) * X1S + 8 * q;
        *reinterpret_cast<float4 *>(p) = make_float4(acc[dx][0], acc[dx][1], acc[dx][2], acc[dx][3]);
        *reinterpret_cast<float4 *>(p + 4) = make_float4(acc[dx][4], acc[dx][5], acc[dx][6], acc[dx][7]);
    }
    __syncthreads();
    // 9*8*32 = 2304 outputs, 9 per thread, x fastest -> coalesced 128-byte rows
    float *ob = out + (size_t)b * obs + ((size_t)dyi * kD) * H * W;
#pragma unroll
    for (int m = 0; m < kD; ++m) {
        const int e = tid + m * B_THREADS;
        const int dx = e >> 8, rr = (e >> 5) & 7, col = e & 31;
        const float *p = smem + (dx * TH + rr) * X1S + col;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < B_WARPS; ++w) s += p[w * (kD * TH * X1S)];
        const int gy = y0 + rr, gx = x0 + col;
        if (gy < H && gx < W) ob[((size_t)dx * H + gy) * W + gx] = finish(s, rc, leaky, slope);
    }
}


// ===========================================================================================
// TMA variants (W % 4 == 0).  Same thread <-> data mapping and the same fma_channel as above.
// ===========================================================================================
template <bool LEAKY>
__device__ __forceinline__ float finish_t(float v, float rc, float slope) {
    v *= rc;
    if (LEAKY) v = v < 0.f ? v * slope : v;
    return v;
}

constexpr int TA_THREADS = 32 * (kD + 1);                  // 9 consumer warps + 1 producer warp
constexpr int X1_BYTES = CK * TH * X1S * 4;                // 9216  (box 36 x 8 x 8)
constexpr int TA_X2_BYTES = CK * A_ROWS2 * X2S * 4;        // 22528 (box 44 x 16 x 8)
constexpr int TA_STAGE_BYTES = X1_BYTES + TA_X2_BYTES;     // 31744

// grid = (tiles_x, tiles_y, B)
// Persistent tile kernel.  grid = min(#tiles, 148 * OCC) CTAs; CTA i walks tiles i, i+grid, ...
// The producer lane streams stages continuously ACROSS tiles, so while the consumer warps finish a
// tile (and its stores drain) the next tile's channels are already landing: per-tile start-up
// latency (first TMA round trip, ~1 us) is paid once per CTA instead of once per tile.
//   PIPE = false: 96 registers, 2 CTAs/SM (20 warps: the SM's four schedulers get 5 warps each);
//   PIPE = true : software-pipelined operand loads (channel c+1 in flight under the FFMAs of c),
//                 ~130 registers, 1 CTA/SM with a 6-stage ring; used when tiles <= 1.5 per SM.
template <bool PIPE> struct T9Cfg {
    static constexpr int kStages = PIPE ? 6 : 3;
    static constexpr int kMinBlocks = PIPE ? 1 : 2;
    static constexpr int kSmem = kStages * TA_STAGE_BYTES + 2 * kStages * 8;
};

// Store 8 consecutive pixels x 9 dx of one row so that every STG.128 fills whole 32-byte sectors:
// lanes (r,q) and (r^1,q) swap halves, then one instruction writes both 16-byte halves of a sector.
template <bool LEAKY>
__device__ __forceinline__ void store_tile(float (&acc)[kD][8], float *__restrict__ out, int b, int dyi, int y0,
                                           int x0, int r, int q, int H, int W, float rc, float slope, size_t obs) {
    const bool odd = r & 1;
    const int row0 = y0 + (r & ~1), gx = x0 + 8 * q + (odd ? 4 : 0);
    const size_t kstride = (size_t)H * W;
    float *o = out + (size_t)b * obs + (((size_t)dyi * kD) * H + row0) * W + gx;
    const bool c_ok = gx < W, r0_ok = row0 < H, r1_ok = row0 + 1 < H;
#pragma unroll
    for (int dx = 0; dx < kD; ++dx, o += kstride) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = finish_t<LEAKY>(acc[dx][i], rc, slope);
        float snd[4], rcv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) snd[i] = odd ? v[i] : v[4 + i];
#pragma unroll
        for (int i = 0; i < 4; ++i) rcv[i] = __shfl_xor_sync(0xffffffffu, snd[i], 4);
        // even lane: (row0, own low half) then (row0+1, partner's low half)
        // odd  lane: (row0, partner's high half) then (row0+1, own high half)
        const float4 first = odd ? make_float4(rcv[0], rcv[1], rcv[2], rcv[3]) : make_float4(v[0], v[1], v[2], v[3]);
        const float4 second = odd ? make_float4(v[4], v[5], v[6], v[7]) : make_float4(rcv[0], rcv[1], rcv[2], rcv[3]);
        if (c_ok && r0_ok) *reinterpret_cast<float4 *>(o) = first;
        if (c_ok && r1_ok) *reinterpret_cast<float4 *>(o + W) = second;
    }
}

template <bool LEAKY, bool PIPE>
__global__ void __launch_bounds__(TA_THREADS, T9Cfg<PIPE>::kMinBlocks)
corr_fwd_tile9_tma(const __grid_constant__ CUtensorMap m1, const __grid_constant__ CUtensorMap m2,
                   float *__restrict__ out, int C, int H, int W, int tiles_x, int tiles_y, int ntiles, float rc,
                   float slope, size_t obs) {
    pdl_wait();
    constexpr int NS = T9Cfg<PIPE>::kStages;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem_raw + NS * TA_STAGE_BYTES);
    uint64_t *empty_bar = full_bar + NS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunk = (C + CK - 1) / CK;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kD);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kD) {  // producer warp: one lane feeds the ring, running ahead across tiles
        if (lane == 0) {
            tma_prefetch_desc(&m1);
            tma_prefetch_desc(&m2);
            int it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
                const int x0 = tx * TW, y0 = ty * TH;
                for (int k = 0; k < nchunk; ++k, ++it) {
                    const int s = it % NS;
                    if (it >= NS) mbar_wait(&empty_bar[s], ((it / NS) - 1) & 1);
                    unsigned char *st = smem_raw + s * TA_STAGE_BYTES;
                    mbar_expect_tx(&full_bar[s], TA_STAGE_BYTES);
                    tma_load_4d(st, &m1, &full_bar[s], x0, y0, k * CK, b);
                    tma_load_4d(st + X1_BYTES, &m2, &full_bar[s], x0 - kMaxDisp, y0 - kMaxDisp, k * CK, b);
                }
            }
        }
        return;
    }

    const int r = lane >> 2, q = lane & 3;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
        float acc[kD][8];
#pragma unroll
        for (int dx = 0; dx < kD; ++dx)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[dx][i] = 0.f;

        for (int k = 0; k < nchunk; ++k, ++it) {
            const int s = it % NS;
            mbar_wait(&full_bar[s], (it / NS) & 1);
            const float *st = reinterpret_cast<const float *>(smem_raw + s * TA_STAGE_BYTES);
            const float *p1 = st + r * X1S + 8 * q;
            const float *p2 = st + CK * TH * X1S + (r + warp) * X2S + 8 * q;
            if (PIPE) {
                ChanOps cur, nxt;
                load_ops(cur, p1, p2);
#pragma unroll
                for (int c = 0; c < CK; ++c) {
                    if (c + 1 < CK) load_ops(nxt, p1 + (c + 1) * TH * X1S, p2 + (c + 1) * A_ROWS2 * X2S);
                    fma_ops(acc, cur);
                    if (c + 1 < CK) cur = nxt;
                }
            } else {
#pragma unroll 2
                for (int c = 0; c < CK; ++c) fma_channel(acc, p1 + c * TH * X1S, p2 + c * A_ROWS2 * X2S);
            }
            release_stage_reads();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
        store_tile<LEAKY>(acc, out, b, warp, ty * TH, tx * TW, r, q, H, W, rc, slope, obs);
    }
}

constexpr int TB_THREADS = 32 * (B_WARPS + 1);             // 8 consumer warps + 1 producer warp
constexpr int TB_NST = 4;
constexpr int TB_X2_BYTES = CK * TH * X2S * 4;             // 11264 (box 44 x 8 x 8)
constexpr int TB_STAGE_BYTES = X1_BYTES + TB_X2_BYTES;     // 20480
constexpr int TB_SMEM = (B_RED * 4 > TB_NST * TB_STAGE_BYTES) ? B_RED * 4 : TB_NST * TB_STAGE_BYTES;

// grid = (tiles_x, tiles_y, B * 9)
template <bool LEAKY>
__global__ void __launch_bounds__(TB_THREADS, 2)
corr_fwd_splitc_tma(const __grid_constant__ CUtensorMap m1, const __grid_constant__ CUtensorMap m2,
                    float *__restrict__ out, int C, int H, int W, float rc, float slope, size_t obs) {
    pdl_wait();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[TB_NST], empty_bar[TB_NST];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int b = blockIdx.z / kD, dyi = blockIdx.z - b * kD;
    const int nchunk = (C + CK - 1) / CK;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TB_NST; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], B_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == B_WARPS) {
        if (lane == 0) {
            tma_prefetch_desc(&m1);
            tma_prefetch_desc(&m2);
            for (int k = 0; k < nchunk; ++k) {
                const int s = k % TB_NST;
                if (k >= TB_NST) mbar_wait(&empty_bar[s], ((k / TB_NST) - 1) & 1);
                unsigned char *st = smem_raw + s * TB_STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], TB_STAGE_BYTES);
                tma_load_4d(st, &m1, &full_bar[s], x0, y0, k * CK, b);
                tma_load_4d(st + X1_BYTES, &m2, &full_bar[s], x0 - kMaxDisp, y0 + dyi - kMaxDisp, k * CK, b);
            }
        }
        return;
    }

    const int r = lane >> 2, q = lane & 3;
    float acc[kD][8];
#pragma unroll
    for (int dx = 0; dx < kD; ++dx)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[dx][i] = 0.f;

    for (int k = 0; k < nchunk; ++k) {
        const int s = k % TB_NST;
        mbar_wait(&full_bar[s], (k / TB_NST) & 1);
        const float *st = reinterpret_cast<const float *>(smem_raw + s * TB_STAGE_BYTES);
        fma_channel(acc, st + (warp * TH + r) * X1S + 8 * q, st + CK * TH * X1S + (warp * TH + r) * X2S + 8 * q);
        release_stage_reads();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    // all 8 consumer warps are past their last stage read -> reuse the ring for the reduction
    asm volatile("bar.sync 1, %0;\n" ::"n"(32 * B_WARPS) : "memory");
    float *smem = reinterpret_cast<float *>(smem_raw);
    float *red = smem + (size_t)warp * (kD * TH * X1S);
#pragma unroll
    for (int dx = 0; dx < kD; ++dx) {
        float *p = red + (dx * TH + r) * X1S + 8 * q;
        *reinterpret_cast<float4 *>(p) = make_float4(acc[dx][0], acc[dx][1], acc[dx][2], acc[dx][3]);
        *reinterpret_cast<float4 *>(p + 4) = make_float4(acc[dx][4], acc[dx][5], acc[dx][6], acc[dx][7]);
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(32 * B_WARPS) : "memory");
    float *ob = out + (size_t)b * obs + ((size_t)dyi * kD) * H * W;
#pragma unroll
    for (int m = 0; m < kD; ++m) {
        const int e = tid + m * (32 * B_WARPS);
        const int dx = e >> 8, rr = (e >> 5) & 7, col = e & 31;
        const float *p = smem + (dx * TH + rr) * X1S + col;
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < B_WARPS; ++w) sum += p[w * (kD * TH * X1S)];
        const int gy = y0 + rr, gx = x0 + col;
        if (gy < H && gx < W) ob[((size_t)dx * H + gy) * W + gx] = finish_t<LEAKY>(sum, rc, slope);
    }
}

template <bool LEAKY>
cudaError_t launch_tma(bool splitc, const CUtensorMap &m1, const CUtensorMap &m2a, const CUtensorMap &m2b,
                       float *out, int B, int C, int H, int W, int tx, int ty, float rc, float slope,
                       cudaStream_t stream) {
    const size_t obs = corr_out_bstride(H, W);
    if (splitc) {
        cudaError_t once = ensure_smem((const void *)corr_fwd_splitc_tma<LEAKY>, TB_SMEM);
        if (once != cudaSuccess) return once;
        usf_launch(corr_fwd_splitc_tma<LEAKY>, dim3(dim3(tx, ty, B * kD)), dim3(TB_THREADS), TB_SMEM, stream, m1, m2b, out, C, H, W, rc, slope, obs);
    } else {
        const int ntiles = tx * ty * B;
        const int sms = sm_count();
        if (ntiles * 2 <= 3 * sms) {  // at most 1.5 tiles per SM: one pipelined CTA per SM
            cudaError_t once = ensure_smem((const void *)corr_fwd_tile9_tma<LEAKY, true>, T9Cfg<true>::kSmem);
            if (once != cudaSuccess) return once;
            const int grid = ntiles < sms ? ntiles : sms;
            usf_launch((corr_fwd_tile9_tma<LEAKY, true>), dim3(grid), dim3(TA_THREADS), T9Cfg<true>::kSmem, stream, 
                m1, m2a, out, C, H, W, tx, ty, ntiles, rc, slope, obs);
        } else {
            cudaError_t once = ensure_smem((const void *)corr_fwd_tile9_tma<LEAKY, false>, T9Cfg<false>::kSmem);
            if (once != cudaSuccess) return once;
            const int grid = ntiles < 2 * sms ? ntiles : 2 * sms;
            usf_launch((corr_fwd_tile9_tma<LEAKY, false>), dim3(grid), dim3(TA_THREADS), T9Cfg<false>::kSmem, stream, 
                m1, m2a, out, C, H, W, tx, ty, ntiles, rc, slope, obs);
        }
    }
    return cudaGetLastError();
}

}  // namespace

int corr_fwd_direct_f32(const float *, const float *, float *, int, int, int, int, unsigned, float, cudaStream_t);

int corr_fwd_f32(const float *x1, const float *x2, float *out, int B, int C, int H, int W, unsigned flags,
                 float slope, cudaStream_t stream) {
    // small odd-width shapes above 256 pixels per image (whole images <= 256 pixels never get here, api.cu sends
    // them to corr_small.cu): L1/L2 resident -> direct kernel (corr_direct.cu); widths TMA can address go to the
    // tiled kernels even when small
    if ((long)B * H * W <= kDirectMaxPixels && (W % 4) != 0) return corr_fwd_direct_f32(x1, x2, out, B, C, H, W, flags, slope, stream);
    const size_t obs = corr_out_bstride(H, W);
    const bool vec = (W % 4 == 0) && (((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)out) % 16 == 0) && obs % 4 == 0;
    const int tx = (W + TW - 1) / TW, ty = (H + TH - 1) / TH;
    const float rc = 1.0f / (float)C;
    const int leaky = (flags & USF_FLAG_LEAKY_RELU) ? 1 : 0;
    // Few tiles and many channels: split dy over CTAs and channels over warps.
    const long tiles = (long)tx * ty * B;
    // fewer tiles than half the SMs: split dy over CTAs and channels over warps; else persistent tiles
    const bool splitc = tiles * 2 < sm_count() && C >= 48;
    if (vec) {
        // TMA path: tensor maps are encoded on the host per call (pointers differ per call) and travel
        // as __grid_constant__ kernel parameters; nothing is cached between calls (re-entrant).
        CUtensorMap m1, m2;
        if (make_nchw_map(&m1, x1, B, C, H, W, X1S, TH, CK) &&
            make_nchw_map(&m2, x2, B, C, H, W, X2S, splitc ? TH : A_ROWS2, CK)) {
            cudaError_t et = leaky ? launch_tma<true>(splitc, m1, m2, m2, out, B, C, H, W, tx, ty, rc, slope, stream)
                                   : launch_tma<false>(splitc, m1, m2, m2, out, B, C, H, W, tx, ty, rc, slope, stream);
            return note_launch(et);
        }
    }
    cudaError_t e;
    if (splitc) {
        dim3 grid(tx, ty, B * kD);
        if (vec) {
            if (cudaError_t once = ensure_smem((const void *)corr_fwd_splitc<true>, B_SMEM)) return note_launch(once);
            corr_fwd_splitc<true><<<grid, B_THREADS, B_SMEM, stream>>>(x1, x2, out, C, H, W, rc, leaky, slope, obs);
        } else {
            if (cudaError_t once = ensure_smem((const void *)corr_fwd_splitc<false>, B_SMEM)) return note_launch(once);
            corr_fwd_splitc<false><<<grid, B_THREADS, B_SMEM, stream>>>(x1, x2, out, C, H, W, rc, leaky, slope, obs);
        }
    } else {
        dim3 grid(tx, ty, B);
        if (vec) {
            if (cudaError_t once = ensure_smem((const void *)corr_fwd_tile9<true>, A_SMEM)) return note_launch(once);
            corr_fwd_tile9<true><<<grid, A_THREADS, A_SMEM, stream>>>(x1, x2, out, C, H, W, rc, leaky, slope, obs);
        } else {
            if (cudaError_t once = ensure_smem((const void *)corr_fwd_tile9<false>, A_SMEM)) return note_launch(once);
            corr_fwd_tile9<false><<<grid, A_THREADS, A_SMEM, stream>>>(x1, x2, out, C, H, W, rc, leaky, slope, obs);
        }
    }
    e = cudaGetLastError();
    return note_launch(e);
}

}  // namespace usf
