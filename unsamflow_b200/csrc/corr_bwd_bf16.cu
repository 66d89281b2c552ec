// corr_bwd_bf16.cu -- cost-volume backward, bf16 storage / fp32 accumulate, tiled + TMA (sm_100a).
//
// Same decomposition as corr_bwd_gather_tma (corr_bwd.cu; formulas and the gx2 = "same gather with e = -d,
// channel 80-k, shifted G" identity are derived at the top of that file): CTA = 8x32 pixels x 32 channels
// (4 consumer warps x 8 channels + one producer lane), X halo tile resident, G streamed one dy at a time
// through a 2-stage full/empty mbarrier ring, thread = 8 pixels x 8 channels (64 fp32 accumulators), both
// gradients in one launch, atomic-free.  What bf16 storage changes (as in corr_fwd_bf16.cu):
//   * shared memory holds bf16.  TMA's innermost coordinate must be 16-byte aligned = 8 bf16, so halo
//     boxes start at x0-8 (not x0-4): the X box is [8 ch][16 rows][56] and a thread's 16-element window
//     sits at elements [8q+4, 8q+20) of its row: 3 LDS.128, unpacked with one shift/mask per element;
//   * G tile of gx1: [9][8][40], 8 values = 1 LDS.128 per dx.  gx2: halo box [9][8][56] at (x0-8, y0+ey);
//     displacement dxi needs channel slot 8-dxi at elements [8q+4+dxi, +8): one or two LDS.128 and a
//     compile-time selection;
//   * lanes map as (row = lane & 7, strip = lane >> 3): a quarter-warp reads 8 different rows, and row strides
//     of 7 / 5 sixteen-byte chunks are conflict-free;
//   * outputs are packed to bf16, 16 bytes per (thread, channel).
// 72 KB per CTA; registers (not shared memory) hold it at 2 CTAs/SM.  Requires W % 8 == 0 and 16-byte aligned
// bases; other shapes fall back (direct kernels / host-side widening).
#include "common.cuh"
#include "tma.cuh"

namespace usf {
namespace {

constexpr int TH = 8, TW = 32;
constexpr int CT = 8, NCW = 4, CG = CT * NCW;           // channels per thread / consumer warps / channels per CTA
constexpr int XR = TH + 2 * kMaxDisp;                   // 16 halo rows
constexpr int XE = 56;                                  // X / gx2-G row: 56 bf16 = 7 chunks, box starts at x0 - 8
constexpr int GE = 40;                                  // gx1 G row: 40 bf16 = 5 chunks
constexpr int XW_BYTES = CT * XR * XE * 2;              // 14336: one warp's 8-channel halo box
constexpr int G1_BYTES = kD * TH * GE * 2;              // 5760
constexpr int G2_BYTES = kD * TH * XE * 2;              // 8064
constexpr int NBAR = NCW + 4;                           // xfull[4], gfull[2], gempty[2]
constexpr int SMEM = NCW * XW_BYTES + 2 * G2_BYTES + NBAR * 8;
constexpr int THREADS = 32 * (NCW + 1);

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);  // .x = lo (low 16 bits)
    return *reinterpret_cast<const uint32_t *>(&p);
}
__device__ __forceinline__ void unpack8(float *dst, const uint4 &v) {
    dst[0] = bf_lo(v.x); dst[1] = bf_hi(v.x); dst[2] = bf_lo(v.y); dst[3] = bf_hi(v.y);
    dst[4] = bf_lo(v.z); dst[5] = bf_hi(v.z); dst[6] = bf_lo(v.w); dst[7] = bf_hi(v.w);
}

// eight G values of displacement DXI for this thread's pixels
template <bool GX2, int DXI>
__device__ __forceinline__ void load_g(float (&g)[8], const unsigned char *__restrict__ stage, int r, int q) {
    if (!GX2) {
        const uint4 v = *reinterpret_cast<const uint4 *>(stage + ((DXI * TH + r) * GE + 8 * q) * 2);
        unpack8(g, v);
    } else {
        // channel slot 8-DXI, elements [8q + 4 + DXI, +8) of the row (box starts at x0 - 8)
        constexpr int S0 = 4 + DXI, A = (S0 / 8) * 8, O = S0 - A;
        const unsigned char *p = stage + (((kD - 1 - DXI) * TH + r) * XE + 8 * q + A) * 2;
        float v[16];
        unpack8(v, *reinterpret_cast<const uint4 *>(p));
        if (O != 0) unpack8(v + 8, *reinterpret_cast<const uint4 *>(p + 16));
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = v[O + i];
    }
}
template <bool GX2, int DXI>
struct GLoader {
    static __device__ __forceinline__ void run(float (&g)[kD][8], const unsigned char *stage, int r, int q) {
        load_g<GX2, DXI>(g[DXI], stage, r, q);
        GLoader<GX2, DXI + 1>::run(g, stage, r, q);
    }
};
template <bool GX2>
struct GLoader<GX2, kD> {
    static __device__ __forceinline__ void run(float (&)[kD][8], const unsigned char *, int, int) {}
};

template <bool GX2>
__device__ __forceinline__ void bwd_tile(unsigned char *smem_raw, const CUtensorMap &mx, const CUtensorMap &mg,
                                         bf16 *__restrict__ gx, int C, int H, int W, int ngroups, int zidx,
                                         float rc) {
    constexpr int GB = GX2 ? G2_BYTES : G1_BYTES;
    unsigned char *gbase = smem_raw + NCW * XW_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(gbase + 2 * G2_BYTES);
    uint64_t *xfull = bars, *gfull = bars + NCW, *gempty = bars + NCW + 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
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
                    tma_load_4d(gbase + s * G2_BYTES, &mg, &gfull[s], x0, y0, dyi * kD, b);
                else
                    tma_load_4d(gbase + s * G2_BYTES, &mg, &gfull[s], x0 - 8, y0 + dyi - kMaxDisp,
                                (kD - 1 - dyi) * kD, b);
                if (dyi == 0) {
#pragma unroll 1
                    for (int w = 0; w < NCW; ++w) {
                        mbar_expect_tx(&xfull[w], XW_BYTES);
                        tma_load_4d(smem_raw + w * XW_BYTES, &mx, &xfull[w], x0 - 8, y0 - kMaxDisp, c0 + w * CT, b);
                    }
                }
            }
        }
        return;
    }

    const int r = lane & 7, q = lane >> 3;
    float acc[CT][8];
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c][i] = 0.f;

    const unsigned char *xw = smem_raw + warp * XW_BYTES + (r * XE + 8 * q) * 2;
    mbar_wait(&xfull[warp], 0);
#pragma unroll 1
    for (int dyi = 0; dyi < kD; ++dyi) {
        const int s = dyi & 1;
        mbar_wait(&gfull[s], (dyi >> 1) & 1);
        float g[kD][8];
        GLoader<GX2, 0>::run(g, gbase + s * G2_BYTES, r, q);
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const unsigned char *p = xw + ((c * XR + dyi) * XE) * 2;
            float v[24];
            unpack8(v, *reinterpret_cast<const uint4 *>(p));
            unpack8(v + 8, *reinterpret_cast<const uint4 *>(p + 16));
            unpack8(v + 16, *reinterpret_cast<const uint4 *>(p + 32));
            // window = elements 4..19 of the 24 loaded (columns x0+8q-4 .. x0+8q+11)
#pragma unroll
            for (int dx = 0; dx < kD; ++dx)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[c][i] = fmaf(g[dx][i], v[4 + dx + i], acc[c][i]);
            if (c == 3) {  // release the G stage: fence + arrive, placed as in corr_bwd.cu (kRelAt)
                release_stage_reads();
                if (lane == 0) mbar_arrive(&gempty[s]);
            }
        }
    }

    const int gy = y0 + r, gxx = x0 + 8 * q;
    if (gy >= H || gxx >= W) return;  // W % 8 == 0: an 8-pixel group is entirely inside or outside
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        const int cc = c0 + warp * CT + c;
        if (cc >= C) break;
        bf16 *o = gx + (((size_t)b * C + cc) * H + gy) * W + gxx;
        *reinterpret_cast<uint4 *>(o) =
            make_uint4(pack_bf16(acc[c][0] * rc, acc[c][1] * rc), pack_bf16(acc[c][2] * rc, acc[c][3] * rc),
                       pack_bf16(acc[c][4] * rc, acc[c][5] * rc), pack_bf16(acc[c][6] * rc, acc[c][7] * rc));
    }
}

// z layout: [0, nz) -> first requested gradient, [nz, 2 nz) -> gx2 when both are requested
__global__ void __launch_bounds__(THREADS, 2)
corr_bwd_gather_tma_bf16(const __grid_constant__ CUtensorMap mx1, const __grid_constant__ CUtensorMap mg1,
                         const __grid_constant__ CUtensorMap mx2, const __grid_constant__ CUtensorMap mg2,
                         bf16 *__restrict__ gx1, bf16 *__restrict__ gx2, int C, int H, int W, int ngroups, int nz,
                         float rc) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    int z = blockIdx.z;
    const bool second = (gx1 == nullptr) || z >= nz;
    if (z >= nz) z -= nz;
    if (second)
        bwd_tile<true>(smem_raw, mx2, mg2, gx2, C, H, W, ngroups, z, rc);
    else
        bwd_tile<false>(smem_raw, mx1, mg1, gx1, C, H, W, ngroups, z, rc);
}

}  // namespace

// Returns USF_ERR_UNSUPPORTED when the shape cannot take the TMA path (caller falls back).
int corr_bwd_bf16_tiled(const void *x1, const void *x2, const void *gout, void *gx1, void *gx2, int B, int C, int H,
                        int W, cudaStream_t stream) {
    const uintptr_t bits = (uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)gout | (uintptr_t)gx1 | (uintptr_t)gx2;
    if ((W % 8) != 0 || (bits & 15)) return USF_ERR_UNSUPPORTED;
    CUtensorMap mx1 = {}, mg1 = {}, mx2 = {}, mg2 = {};
    bool ok = true;
    if (gx1) ok = make_nchw_map_bf16(&mx1, x2, B, C, H, W, XE, XR, CT) && make_nchw_map_bf16(&mg1, gout, B, kK, H, W, GE, TH, kD);
    if (ok && gx2) ok = make_nchw_map_bf16(&mx2, x1, B, C, H, W, XE, XR, CT) && make_nchw_map_bf16(&mg2, gout, B, kK, H, W, XE, TH, kD);
    if (!ok) return USF_ERR_UNSUPPORTED;
    if (!gx1) { mx1 = mx2; mg1 = mg2; }
    if (!gx2) { mx2 = mx1; mg2 = mg1; }
    if (cudaError_t once = ensure_smem((const void *)corr_bwd_gather_tma_bf16, SMEM)) return note_launch(once);
    const int ngroups = (C + CG - 1) / CG, nz = B * ngroups;
    dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, nz * ((gx1 && gx2) ? 2 : 1));
    corr_bwd_gather_tma_bf16<<<grid, THREADS, SMEM, stream>>>(mx1, mg1, mx2, mg2, (bf16 *)gx1, (bf16 *)gx2, C, H, W,
                                                              ngroups, nz, 1.0f / (float)C);
    return note_launch(cudaGetLastError());
}

}  // namespace usf
