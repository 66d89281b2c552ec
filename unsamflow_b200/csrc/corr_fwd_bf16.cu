// corr_fwd_bf16.cu -- cost-volume forward, bf16 storage / fp32 accumulate, tiled + TMA (sm_100a).
//
// Same decomposition as corr_fwd_tile9_tma (corr_fwd.cu): persistent CTAs, 8x32 pixel tiles, 9 consumer
// warps = 9 dy, one producer lane streaming 8-channel stages through full/empty mbarriers, a thread owns
// 8 pixels x 9 dx (72 fp32 accumulators).  What bf16 storage changes:
//   * shared memory holds bf16: per channel a thread issues 1 + 3 LDS.128 (x1 strip; 24-element x2 window)
//     instead of 2 + 4 -- 16 wavefronts instead of 24 on the pipe that bounds the fp32 kernel -- and unpacks
//     with one shift/mask per element;
//   * TMA's innermost coordinate must be 16-byte aligned = 8 bf16, so the x2 halo box starts at x0-8 (not
//     x0-4) and the thread's 16-element window sits at elements [4,20) of the 24 it loads (static indices);
//   * row strides are 5 (x1: 40 elements) and 7 (x2: 56 elements) 16-byte chunks and lanes map as
//     (row = lane & 7, column group = lane >> 3): a quarter-warp reads 8 different rows -> conflict-free;
//   * outputs are packed to bf16 and each lane stores 16 bytes per dx; a warp store covers 8 rows x 64
//     contiguous bytes (whole sectors, no shuffle needed).
// Requires W % 8 == 0 (global strides multiple of 16 bytes); other shapes use the direct kernels / the
// host-side widening path.
#include "common.cuh"
#include "tma.cuh"

namespace usf {
namespace {

constexpr int TH = 8, TW = 32, CK = 8;
constexpr int X1E = 40;                       // x1 row: 40 bf16 = 5 chunks
constexpr int X2E = 56;                       // x2 row: 56 bf16 = 7 chunks, box starts at x0 - 8
constexpr int X2R = TH + 2 * kMaxDisp;        // 16 halo rows
constexpr int X1_BYTES = CK * TH * X1E * 2;   // 5120
constexpr int X2_BYTES = CK * X2R * X2E * 2;  // 14336
constexpr int STAGE_BYTES = X1_BYTES + X2_BYTES;  // 19456 = 152 * 128
constexpr int NS = 4;
constexpr int THREADS = 32 * (kD + 1);
constexpr int SMEM = NS * STAGE_BYTES + 2 * NS * 8;

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);  // .x = lo (low 16 bits), .y = hi
    return *reinterpret_cast<const uint32_t *>(&p);
}

template <bool LEAKY>
__global__ void __launch_bounds__(THREADS, 2)
corr_fwd_tile9_tma_bf16(const __grid_constant__ CUtensorMap m1, const __grid_constant__ CUtensorMap m2,
                        bf16 *__restrict__ out, int C, int H, int W, int tiles_x, int tiles_y, int ntiles, float rc,
                        float slope, size_t obs) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem_raw + NS * STAGE_BYTES);
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

    if (warp == kD) {
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
                    unsigned char *st = smem_raw + s * STAGE_BYTES;
                    mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                    tma_load_4d(st, &m1, &full_bar[s], x0, y0, k * CK, b);
                    tma_load_4d(st + X1_BYTES, &m2, &full_bar[s], x0 - 8, y0 - kMaxDisp, k * CK, b);
                }
            }
        }
        return;
    }

    const int r = lane & 7, q = lane >> 3;
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
            const unsigned char *st = smem_raw + s * STAGE_BYTES;
            const unsigned char *p1 = st + r * (X1E * 2) + q * 16;
            const unsigned char *p2 = st + X1_BYTES + (r + warp) * (X2E * 2) + q * 16;
#pragma unroll 2
            for (int c = 0; c < CK; ++c) {
                const uint4 a4 = *reinterpret_cast<const uint4 *>(p1 + c * (TH * X1E * 2));
                const unsigned char *pw = p2 + c * (X2R * X2E * 2);
                const uint4 w0 = *reinterpret_cast<const uint4 *>(pw);
                const uint4 w1 = *reinterpret_cast<const uint4 *>(pw + 16);
                const uint4 w2 = *reinterpret_cast<const uint4 *>(pw + 32);
                const float a[8] = {bf_lo(a4.x), bf_hi(a4.x), bf_lo(a4.y), bf_hi(a4.y),
                                    bf_lo(a4.z), bf_hi(a4.z), bf_lo(a4.w), bf_hi(a4.w)};
                // elements 4..19 of the 24 loaded: words 2..9 = w0.z, w0.w, w1.x .. w1.w, w2.x, w2.y
                const float e[16] = {bf_lo(w0.z), bf_hi(w0.z), bf_lo(w0.w), bf_hi(w0.w), bf_lo(w1.x), bf_hi(w1.x),
                                     bf_lo(w1.y), bf_hi(w1.y), bf_lo(w1.z), bf_hi(w1.z), bf_lo(w1.w), bf_hi(w1.w),
                                     bf_lo(w2.x), bf_hi(w2.x), bf_lo(w2.y), bf_hi(w2.y)};
#pragma unroll
                for (int dx = 0; dx < kD; ++dx)
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[dx][i] = fmaf(a[i], e[dx + i], acc[dx][i]);
            }
            release_stage_reads();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }

        const int gy = ty * TH + r, gx = tx * TW + 8 * q;
        if (gy < H && gx < W) {  // W % 8 == 0: an 8-pixel group is entirely inside or outside
            bf16 *o = out + (size_t)b * obs + (((size_t)warp * kD) * H + gy) * W + gx;
            const size_t kstride = (size_t)H * W;
#pragma unroll
            for (int dx = 0; dx < kD; ++dx, o += kstride) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    v[i] = acc[dx][i] * rc;
                    if (LEAKY) v[i] = v[i] < 0.f ? v[i] * slope : v[i];
                }
                *reinterpret_cast<uint4 *>(o) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                                                           pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            }
        }
    }
}

}  // namespace

// Returns USF_ERR_UNSUPPORTED when the shape cannot take the TMA path (caller falls back).
int corr_fwd_bf16_tiled(const void *x1, const void *x2, void *out, int B, int C, int H, int W, unsigned flags,
                        float slope, cudaStream_t stream) {
    const size_t obs = corr_out_bstride(H, W);
    if ((W % 8) != 0 || (((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)out) & 15) || obs % 8 != 0) return USF_ERR_UNSUPPORTED;
    CUtensorMap m1, m2;
    if (!make_nchw_map_bf16(&m1, x1, B, C, H, W, X1E, TH, CK) || !make_nchw_map_bf16(&m2, x2, B, C, H, W, X2E, X2R, CK))
        return USF_ERR_UNSUPPORTED;
    const int tx = (W + TW - 1) / TW, ty = (H + TH - 1) / TH, ntiles = tx * ty * B;
    const int grid = ntiles < 2 * sm_count() ? ntiles : 2 * sm_count();
    const float rc = 1.0f / (float)C;
    cudaError_t e;
    if (flags & USF_FLAG_LEAKY_RELU) {
        if ((e = ensure_smem((const void *)corr_fwd_tile9_tma_bf16<true>, SMEM)) != cudaSuccess) return note_launch(e);
        corr_fwd_tile9_tma_bf16<true><<<grid, THREADS, SMEM, stream>>>(m1, m2, (bf16 *)out, C, H, W, tx, ty, ntiles, rc, slope, obs);
    } else {
        if ((e = ensure_smem((const void *)corr_fwd_tile9_tma_bf16<false>, SMEM)) != cudaSuccess) return note_launch(e);
        corr_fwd_tile9_tma_bf16<false><<<grid, THREADS, SMEM, stream>>>(m1, m2, (bf16 *)out, C, H, W, tx, ty, ntiles, rc, slope, obs);
    }
    return note_launch(cudaGetLastError());
}

}  // namespace usf
