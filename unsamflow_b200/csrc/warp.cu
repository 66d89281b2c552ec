// warp.cu -- flow_warp (bilinear backward warp) forward + backward for sm_100a.
//
// Replaces: utils/warp_utils.py:97-106 flow_warp = mesh_grid (:7-13, built on the CPU and copied
// to the device on every call) + norm_grid (:16-23) + F.grid_sample(bilinear, align_corners=True,
// padding border|zeros) and its autograd (ATen grid_sampler_2d_backward: fastAtomicAdd scatter).
//
// Coordinates are computed in registers with the reference's exact fp32 operation order
// (no FMA contraction; SURVEY.md 7 hard part 3):
//     v = float(i) + flow;  g = 2*v/(size-1) - 1;  ix = ((g + 1)/2) * (size-1);  [border: clip]
//
// Kernels (HBM/L2-bound gathers; one thread = one pixel x one channel slice, pixels consecutive
// across a warp so every access is coalesced or near-coalesced):
//   warp_fwd_kernel     4 predicated tap loads + 1 store per channel.
//   warp_bwd_kernel     g_flow: per-slice partial sums over channels, fixed-order reduction through
//                       shared memory.  Optionally (a) emits the tap lists used by the atomic-free
//                       g_x gather below, or (b) scatters g_x with red.global.add.f32 (A/B mode).
//   warp_gx_gather      atomic-free g_x: every TARGET pixel owns a list of (source pixel, weight)
//                       pairs (built once per call, channel independent, sorted by source so the
//                       sum order is deterministic) and gathers  g_x[c,t] = sum_i w_i * gO[c,src_i]
//                       with plain loads and one plain store: no float atomics, no zero-fill.
//                       Targets hit by more than kSlots taps (e.g. a region clipped onto the image
//                       border) keep their first kSlots taps in the list; the rest go through a side
//                       list handled with float atomics (warp_gx_overflow) after the gather has written g_x.
#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "warp_coord.cuh"

namespace usf {
namespace {

constexpr int kBlock = 256;
constexpr int kSlots = 8;  // list capacity per target pixel
constexpr int kSlotsS = 16;  // ... in the shared-memory lists of the fused small-image kernel


// ---------------------------------------------------------------------------------------------
// Persistent over pixel blocks (grid = resident CTAs when the level has more blocks than that): the flow of
// the NEXT block is requested before the taps of the current one are consumed, so a thread's life is one
// memory round trip per pixel instead of two back-to-back (flow, then taps).  With cold inputs -- the normal
// case inside a training step, whose working set is far larger than L2 -- that serialised pair was the
// limiter (1.5 TB/s on the 256x832 loss-side warp).
// CC > 0: channel count known at compile time (the 3-channel loss-side warp): the channel loop unrolls
// completely and all 12 taps are in flight at once; with a run-time C = 3 the unroll-by-4 loop degenerates to
// its one-channel-at-a-time remainder, i.e. three dependent round trips.
template <int NS, typename T, int CC = 0>
__global__ void __launch_bounds__(kBlock, CC ? 6 : 5)
warp_fwd_kernel(const T *__restrict__ x, const float *__restrict__ flow, T *__restrict__ out, int N,
                int Crt, int H, int W, int border, int rcp, FastDiv dHW, FastDiv dW, AxisK kx, AxisK ky) {
    pdl_wait();
    const int C = CC ? CC : Crt;
    constexpr int PPB = kBlock / NS;
    constexpr int UN = (NS >= 4 && NS <= 8) ? 8 : 4;  // channels in flight per thread (16+ slices: 4-6 channels each)
    const int pl = threadIdx.x % PPB, cs = threadIdx.x / PPB;
    const int HW = H * W, stride = gridDim.x * PPB;
    int p = blockIdx.x * PPB + pl;
    float u = 0.f, v = 0.f;
    if (p < N) {
        const int b = dHW.div(p), rem = p - b * HW;
        u = __ldg(flow + (size_t)b * 2 * HW + rem);
        v = __ldg(flow + (size_t)b * 2 * HW + HW + rem);
    }
    while (p < N) {
        const int b = dHW.div(p), rem = p - b * HW;
        const int y = dW.div(rem), xo = rem - y * W;
        const int pn = p + stride;
        float un = 0.f, vn = 0.f;
        if (pn < N) {
            const int bn = dHW.div(pn), remn = pn - bn * HW;
            un = __ldg(flow + (size_t)bn * 2 * HW + remn);
            vn = __ldg(flow + (size_t)bn * 2 * HW + HW + remn);
        }
        const Taps t = make_taps_k((float)y, (float)xo, u, v, H, W, kx, ky, border, rcp);
        const float nw = t.ex * t.ey, ne = t.wx * t.ey, sw = t.ex * t.wy, se = t.wx * t.wy;
        // two row pointers (taps nw/ne and sw/se are neighbours) bumped by one slice stride per channel:
        // no per-tap address arithmetic in the loop; out-of-image taps are predicated off, never read
        const size_t step = (size_t)NS * HW;
        const T *p0 = x + ((size_t)b * C + cs) * HW + t.o_nw;
        const T *p1 = p0 + W;
        T *o = out + ((size_t)b * C + cs) * HW + rem;
#pragma unroll(CC ? CC : UN)
        for (int c = cs; c < C; c += NS, p0 += step, p1 += step, o += step) {
            float acc = (t.k_nw ? ldf(p0) : 0.f) * nw;
            acc = fmaf(t.k_ne ? ldf(p0 + 1) : 0.f, ne, acc);
            acc = fmaf(t.k_sw ? ldf(p1) : 0.f, sw, acc);
            acc = fmaf(t.k_se ? ldf(p1 + 1) : 0.f, se, acc);
            stf(o, acc);
        }
        p = pn;
        u = un;
        v = vn;
    }
}

// ---------------------------------------------------------------------------------------------
// The loss-side calls (losses/flow_loss.py:130-131: a 3-channel image, gradient wrt the flow only) are ISSUE-bound,
// not memory-bound (69 % issue-active with DRAM at 22-26 %, ncu round 2), and 40 % of their instructions were 64-bit
// address arithmetic.  These two kernels are the C = 3 forms of warp_fwd_kernel / warp_bwd_kernel<.., 0> with
//   * 32-bit element offsets (api.cu bounds every tensor of a call to 2^31-1 elements): one IMAD.WIDE per row
//     pointer instead of 64-bit multiply chains;
//   * the (image, pixel) split of the NEXT pixel, needed for the flow prefetch, carried into the next iteration
//     instead of being recomputed.
// Same arithmetic, bit-identical results.  (The multi-channel kernels keep their pointer-bump loops: there the
// offsets cost an IMAD.WIDE per load and a few registers, and the same change made them 15-30 % slower.)
// ---------------------------------------------------------------------------------------------
template <typename T, bool BWD>
__global__ void __launch_bounds__(kBlock, 6)
warp_rgb_kernel(const T *__restrict__ x, const float *__restrict__ flow, const T *__restrict__ gout, T *__restrict__ out,
                float *__restrict__ gflow, int N, int H, int W, int border, int rcp, FastDiv dHW, FastDiv dW, AxisK kx,
                AxisK ky) {
    pdl_wait();
    constexpr int C = 3;
    const int HW = H * W, stride = gridDim.x * kBlock;
    int p = blockIdx.x * kBlock + threadIdx.x;
    float u = 0.f, v = 0.f;
    int b = 0, rem = 0;
    if (p < N) {
        b = dHW.div(p);
        rem = p - b * HW;
        const int fo = b * 2 * HW + rem;
        u = __ldg(flow + fo);
        v = __ldg(flow + fo + HW);
    }
    while (p < N) {
        const int y = dW.div(rem), xo = rem - y * W;
        const int pn = p + stride;
        float un = 0.f, vn = 0.f;
        int bn = 0, remn = 0;
        if (pn < N) {
            bn = dHW.div(pn);
            remn = pn - bn * HW;
            const int fo = bn * 2 * HW + remn;
            un = __ldg(flow + fo);
            vn = __ldg(flow + fo + HW);
        }
        const Taps t = make_taps_k((float)y, (float)xo, u, v, H, W, kx, ky, border, rcp);
        const int plane = b * C * HW;
        int i0 = plane + t.o_nw, io = plane + rem;
        if (!BWD) {
            const float nw = t.ex * t.ey, ne = t.wx * t.ey, sw = t.ex * t.wy, se = t.wx * t.wy;
#pragma unroll
            for (int c = 0; c < C; ++c, i0 += HW, io += HW) {
                const T *p0 = x + i0, *p1 = x + (i0 + W);
                float acc = (t.k_nw ? ldf(p0) : 0.f) * nw;
                acc = fmaf(t.k_ne ? ldf(p0 + 1) : 0.f, ne, acc);
                acc = fmaf(t.k_sw ? ldf(p1) : 0.f, sw, acc);
                acc = fmaf(t.k_se ? ldf(p1 + 1) : 0.f, se, acc);
                stf(out + io, acc);
            }
        } else {
            float gix = 0.f, giy = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c, i0 += HW, io += HW) {
                const T *p0 = x + i0, *p1 = x + (i0 + W);
                const float go = ldf(gout + io);
                const float vnw = t.k_nw ? ldf(p0) : 0.f, vne = t.k_ne ? ldf(p0 + 1) : 0.f;
                const float vsw = t.k_sw ? ldf(p1) : 0.f, vse = t.k_se ? ldf(p1 + 1) : 0.f;
                gix = fmaf(go, fmaf(t.ey, vne - vnw, t.wy * (vse - vsw)), gix);
                giy = fmaf(go, fmaf(t.ex, vsw - vnw, t.wx * (vse - vne)), giy);
            }
            // grid_sampler backward: ggrid = mult * gi; norm_grid backward: (g / (size-1)) * 2
            const float gxn = __fmul_rn(t.mx, gix), gyn = __fmul_rn(t.my, giy);
            const float dfx = rcp ? __fmul_rn(gxn, kx.rcp) : div_by_size(gxn, kx);
            const float dfy = rcp ? __fmul_rn(gyn, ky.rcp) : div_by_size(gyn, ky);
            const int fo = b * 2 * HW + rem;
            gflow[fo] = __fmul_rn(dfx, 2.0f);
            gflow[fo + HW] = __fmul_rn(dfy, 2.0f);
        }
        p = pn;
        b = bn;
        rem = remn;
        u = un;
        v = vn;
    }
}

// ---------------------------------------------------------------------------------------------
// Workspace of the atomic-free g_x path, N = B*H*W target pixels:
//   int   cnt[N]            taps received per target
//   int   ovf_n[4]          [0] number of overflow taps, [1] CTAs of the gather that have finished
//   (cnt and ovf_n must be zero when the tap kernel starts: a memset node per call, or -- with
//   USF_FLAG_WORKSPACE_CLEAN -- the caller's promise; the gather leaves both zeroed again)
//   int2  ent[N][kSlots]    (source pixel inside its image plane, weight bits)
//   int   ovf_tgt[4N], ovf_src[4N]; float ovf_w[4N]     overflow side list (worst case: every tap)
struct Workspace {
    int *cnt;
    int *ovf_n;
    int2 *ent;
    int *ovf_tgt;
    int *ovf_src;
    float *ovf_w;
};
// cnt is padded to a multiple of 16 bytes so that ent (read as int4) stays 16-byte aligned
__host__ __device__ inline size_t cnt_bytes(size_t N) { return (N * 4 + 15) / 16 * 16; }
__host__ __device__ inline size_t ws_bytes(size_t N) {
    return cnt_bytes(N) + 16 + N * kSlots * 8 + 3 * (4 * N) * 4;
}
inline Workspace carve(void *base, size_t N) {
    char *p = (char *)base;
    Workspace w;
    w.cnt = (int *)p;               p += cnt_bytes(N);
    w.ovf_n = (int *)p;             p += 16;
    w.ent = (int2 *)p;              p += N * kSlots * 8;
    w.ovf_tgt = (int *)p;           p += 4 * N * 4;
    w.ovf_src = (int *)p;           p += 4 * N * 4;
    w.ovf_w = (float *)p;
    return w;
}

__device__ __forceinline__ void emit_tap(const Workspace &ws, int plane0, int off, bool ok, float w,
                                         int src) {
    if (!ok || w == 0.f) return;
    const int tgt = plane0 + off;
    const int slot = atomicAdd(ws.cnt + tgt, 1);  // integer bookkeeping only; sums stay atomic-free
    if (slot < kSlots) {
        ws.ent[(size_t)tgt * kSlots + slot] = make_int2(src, __float_as_int(w));
    } else {
        const int j = atomicAdd(ws.ovf_n, 1);
        ws.ovf_tgt[j] = tgt;
        ws.ovf_src[j] = src;
        ws.ovf_w[j] = w;
    }
}

// MODE 0: g_flow only.  MODE 1: (g_flow) + emit tap lists.  MODE 2: (g_flow) + atomic g_x scatter.
template <int NS, int MODE, typename T, int CC = 0>
__global__ void __launch_bounds__(kBlock, (NS == 1 && !CC) ? 4 : 5)  // (5 would spill the generic single-slice form)
warp_bwd_kernel(const T *__restrict__ x, const float *__restrict__ flow, const T *__restrict__ gout,
                T *__restrict__ gx, float *__restrict__ gflow, Workspace ws, int N, int Crt, int H, int W,
                int border, int rcp, FastDiv dHW, FastDiv dW, AxisK kx, AxisK ky) {
    pdl_wait();
    const int C = CC ? CC : Crt;  // see warp_fwd_kernel
    constexpr int PPB = kBlock / NS;
    constexpr int UN = (NS >= 4 && NS <= 8) ? 8 : 4;  // channels in flight per thread (16+ slices: 4-6 channels each)
    __shared__ float red[2][NS][PPB];
    const int pl = threadIdx.x % PPB, cs = threadIdx.x / PPB;
    const int HW = H * W, stride = gridDim.x * PPB;
    // persistent over pixel blocks, flow of the next block prefetched (see warp_fwd_kernel); the loop bound
    // is uniform over the CTA because of the barriers in the reduction
    int p = blockIdx.x * PPB + pl;
    float u = 0.f, v = 0.f;
    if (p < N) {
        const int b = dHW.div(p), rem = p - b * HW;
        u = __ldg(flow + (size_t)b * 2 * HW + rem);
        v = __ldg(flow + (size_t)b * 2 * HW + HW + rem);
    }
    for (int base = blockIdx.x * PPB; base < N; base += stride, p += stride) {
        const bool live = p < N;
        const int pn = p + stride;
        float un = 0.f, vn = 0.f;
        if (pn < N) {
            const int bn = dHW.div(pn), remn = pn - bn * HW;
            un = __ldg(flow + (size_t)bn * 2 * HW + remn);
            vn = __ldg(flow + (size_t)bn * 2 * HW + HW + remn);
        }
        float gix = 0.f, giy = 0.f;
        int b = 0, rem = 0;
        Taps t;
        if (live) {
            b = dHW.div(p);
            rem = p - b * HW;
            const int y = dW.div(rem), xo = rem - y * W;
            t = make_taps_k((float)y, (float)xo, u, v, H, W, kx, ky, border, rcp);
            const float nw = t.ex * t.ey, ne = t.wx * t.ey, sw = t.ex * t.wy, se = t.wx * t.wy;
            if (MODE == 1) {
                // every slice has computed the same taps: spread the four list appends (an atomic with a
                // dependent store each) over up to four slices instead of leaving them all to slice 0 while the
                // others wait at the barrier below (a third of this kernel's stall samples, ncu)
                constexpr int EM = NS >= 4 ? 4 : NS;  // emitting slices
                if (cs < EM) {
                    if (cs == 0) emit_tap(ws, b * HW, t.o_nw, t.k_nw, nw, rem);
                    if (cs == 1 % EM) emit_tap(ws, b * HW, t.o_ne, t.k_ne, ne, rem);
                    if (cs == 2 % EM) emit_tap(ws, b * HW, t.o_sw, t.k_sw, sw, rem);
                    if (cs == 3 % EM) emit_tap(ws, b * HW, t.o_se, t.k_se, se, rem);
                }
            }
            const size_t step = (size_t)NS * HW;
            const T *gb = gout + ((size_t)b * C + cs) * HW + rem;
            const T *p0 = x + ((size_t)b * C + cs) * HW + t.o_nw;
            const T *p1 = p0 + W;
            T *q0 = (MODE == 2) ? gx + ((size_t)b * C + cs) * HW + t.o_nw : nullptr;
            if (gflow != nullptr || MODE == 2) {
#pragma unroll(CC ? CC : UN)
                for (int c = cs; c < C; c += NS, gb += step, p0 += step, p1 += step) {
                    const float go = ldf(gb);
                    if (gflow != nullptr) {
                        const float vnw = t.k_nw ? ldf(p0) : 0.f, vne = t.k_ne ? ldf(p0 + 1) : 0.f;
                        const float vsw = t.k_sw ? ldf(p1) : 0.f, vse = t.k_se ? ldf(p1 + 1) : 0.f;
                        gix = fmaf(go, fmaf(t.ey, vne - vnw, t.wy * (vse - vsw)), gix);
                        giy = fmaf(go, fmaf(t.ex, vsw - vnw, t.wx * (vse - vne)), giy);
                    }
                    if (MODE == 2) {
                        if (t.k_nw) atomic_addf(q0, nw * go);
                        if (t.k_ne) atomic_addf(q0 + 1, ne * go);
                        if (t.k_sw) atomic_addf(q0 + W, sw * go);
                        if (t.k_se) atomic_addf(q0 + W + 1, se * go);
                        q0 += step;
                    }
                }
            }
        }
        u = un;
        v = vn;
        if (gflow == nullptr) continue;
        bool writer = live;
        if (NS > 1) {
            red[0][cs][pl] = gix;
            red[1][cs][pl] = giy;
            __syncthreads();
            writer = live && cs == 0;
            if (writer) {
                gix = red[0][0][pl];
                giy = red[1][0][pl];
#pragma unroll
                for (int s = 1; s < NS; ++s) {
                    gix += red[0][s][pl];
                    giy += red[1][s][pl];
                }
            }
            __syncthreads();  // red is rewritten by the next pixel block
        }
        if (writer) {
            // grid_sampler backward: ggrid = mult * gi; norm_grid backward: (g / (size-1)) * 2
            const float gxn = __fmul_rn(t.mx, gix), gyn = __fmul_rn(t.my, giy);
            const float dfx = rcp ? __fmul_rn(gxn, kx.rcp) : div_by_size(gxn, kx);
            const float dfy = rcp ? __fmul_rn(gyn, ky.rcp) : div_by_size(gyn, ky);
            gflow[(size_t)b * 2 * HW + rem] = __fmul_rn(dfx, 2.0f);
            gflow[(size_t)b * 2 * HW + HW + rem] = __fmul_rn(dfy, 2.0f);
        }
    }
}

__device__ __forceinline__ void cswap(int &ka, float &wa, int &kb, float &wb) {
    if (kb < ka) {
        const int tk = ka; ka = kb; kb = tk;
        const float tw = wa; wa = wb; wb = tw;
    }
}

template <int NS, typename T>
// register cap: left alone, the NS = 1 form takes 128 registers and runs at a quarter of the occupancy
__global__ void __launch_bounds__(kBlock, NS == 1 ? 3 : 5)
warp_gx_gather(const T *__restrict__ gout, T *__restrict__ gx, Workspace ws, int N, int C, int HW, FastDiv dHW) {
    pdl_wait();
    constexpr int PPB = kBlock / NS;
    // slice 0 of every pixel loads and sorts the tap list once; the other slices read it from here
    __shared__ int s_src[NS > 1 ? kSlots : 1][NS > 1 ? PPB : 1];
    __shared__ float s_w[NS > 1 ? kSlots : 1][NS > 1 ? PPB : 1];
    __shared__ int s_n[NS > 1 ? PPB : 1];
    const int pl = threadIdx.x % PPB;
    const int p = blockIdx.x * PPB + pl;
    const int cs = threadIdx.x / PPB;
    const bool live = p < N;
    int n = 0;
    int src[kSlots];
    float w[kSlots];
    if (cs == 0) {
        // the first two entry pairs are requested together with the counter (no dependent round trip for the
        // common case of <= 4 taps; slots past the counter hold stale bytes that are never used)
        const int4 *e4 = reinterpret_cast<const int4 *>(ws.ent + (size_t)(live ? p : 0) * kSlots);
        const int4 e01 = e4[0], e23 = e4[1];
        n = live ? min(ws.cnt[p], kSlots) : 0;
        if (live) ws.cnt[p] = 0;  // leave the counters clean for the next call on this workspace
#pragma unroll
        for (int i = 0; i < kSlots / 2; ++i) {
            int4 e = i == 0 ? e01 : (i == 1 ? e23 : make_int4(0, 0, 0, 0));
            if (i >= 2 && 2 * i < n) e = e4[i];
            src[2 * i] = (2 * i < n) ? e.x : 0x7fffffff;
            w[2 * i] = (2 * i < n) ? __int_as_float(e.y) : 0.f;
            src[2 * i + 1] = (2 * i + 1 < n) ? e.z : 0x7fffffff;
            w[2 * i + 1] = (2 * i + 1 < n) ? __int_as_float(e.w) : 0.f;
        }
        // 19-comparator sorting network for 8 keys: list order no longer depends on atomic timing
#define USF_CS(a, b) cswap(src[a], w[a], src[b], w[b])
        USF_CS(0, 1); USF_CS(2, 3); USF_CS(4, 5); USF_CS(6, 7);
        USF_CS(0, 2); USF_CS(1, 3); USF_CS(4, 6); USF_CS(5, 7);
        USF_CS(1, 2); USF_CS(5, 6); USF_CS(0, 4); USF_CS(3, 7);
        USF_CS(1, 5); USF_CS(2, 6);
        USF_CS(1, 4); USF_CS(3, 6);
        USF_CS(2, 4); USF_CS(3, 5);
        USF_CS(3, 4);
#undef USF_CS
        if (NS > 1) {
#pragma unroll
            for (int i = 0; i < kSlots; ++i) {
                s_src[i][pl] = src[i];
                s_w[i][pl] = w[i];
            }
            s_n[pl] = n;
        }
    }
    if (NS > 1) {
        __syncthreads();
        if (cs != 0) {
            n = s_n[pl];
#pragma unroll
            for (int i = 0; i < kSlots; ++i) {
                src[i] = s_src[i][pl];
                w[i] = s_w[i][pl];
            }
        }
    }
    if (live) {
        const int b = dHW.div(p), rem = p - b * HW;
        const size_t step = (size_t)NS * HW;
        const T *plane = gout + ((size_t)b * C + cs) * HW;
        T *o = gx + ((size_t)b * C + cs) * HW + rem;
        // most targets receive <= 4 taps (exactly 4 for a smooth flow): the upper half is a rarely taken branch
        // (unroll 8 here costs registers and occupancy: measured 37 -> 46 us at 8x32x64x208)
#pragma unroll 4
        for (int c = cs; c < C; c += NS, plane += step, o += step) {
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i < n) acc = fmaf(w[i], ldf(plane + src[i]), acc);
            if (n > 4) {
#pragma unroll
                for (int i = 4; i < kSlots; ++i)
                    if (i < n) acc = fmaf(w[i], ldf(plane + src[i]), acc);
            }
            stf(o, acc);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Small images (H*W <= 1024: the coarse pyramid levels): the whole backward in ONE kernel, no workspace.
// The three-kernel path above costs a memset node and three dependent launches (~15 us) for a few microseconds
// of work there.  A CTA owns PPB target pixels of one image.  Every CTA rebuilds the image's taps (H*W
// coordinate computations, cheap at this size) and keeps the ones that land on ITS targets in shared-memory
// lists (shared-memory integer atomics for the slots, sorted by source afterwards: deterministic order);
// taps beyond the 16 list slots go to a shared-memory side list that the target's own threads add before the
// single plain store of each output -- no floating-point atomics at all, no zero-fill, no workspace.
// ---------------------------------------------------------------------------------------------
template <int NS, typename T>
__global__ void __launch_bounds__(kBlock)
warp_bwd_small(const T *__restrict__ x, const float *__restrict__ flow, const T *__restrict__ gout,
               T *__restrict__ gx, float *__restrict__ gflow, int C, int H, int W, int border, int rcp,
               FastDiv dW, FastDiv dBpi, AxisK kx, AxisK ky) {
    pdl_wait();
    constexpr int PPB = kBlock / NS;
    extern __shared__ __align__(16) int sm_i[];
    const int HW = H * W, cap = 4 * HW;
    int *s_cnt = sm_i;                                            // [PPB] taps per target, [PPB] = overflow count
    int *s_src = s_cnt + PPB + 4;                                 // [kSlotsS][PPB]
    float *s_w = reinterpret_cast<float *>(s_src + kSlotsS * PPB); // [kSlotsS][PPB]
    float *red = s_w + kSlotsS * PPB;                              // [2][NS][PPB]
    int *o_tl = reinterpret_cast<int *>(red + 2 * kBlock);        // overflow: local target, source, weight
    int *o_src = o_tl + cap;
    float *o_w = reinterpret_cast<float *>(o_src + cap);
    const int tid = threadIdx.x;
    const int b = dBpi.div(blockIdx.x), t0 = (blockIdx.x - b * dBpi.d) * PPB;
    const float *fl = flow + (size_t)b * 2 * HW;

    for (int i = tid; i <= PPB; i += kBlock) s_cnt[i] = 0;
    __syncthreads();
    // the image's flow first (<= 4 pixels per thread, all requests in flight together), then the taps
    float fu[4], fv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int src = tid + k * kBlock;
        fu[k] = src < HW ? __ldg(fl + src) : 0.f;
        fv[k] = src < HW ? __ldg(fl + HW + src) : 0.f;
    }
    #pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
        const int src = tid + k4 * kBlock;
        if (src >= HW) break;
        const int y = dW.div(src), xo = src - y * W;
        const Taps t = make_taps_k((float)y, (float)xo, fu[k4], fv[k4], H, W, kx, ky, border, rcp);
        const int off[4] = {t.o_nw, t.o_ne, t.o_sw, t.o_se};
        const bool ok[4] = {t.k_nw, t.k_ne, t.k_sw, t.k_se};
        const float wt[4] = {t.ex * t.ey, t.wx * t.ey, t.ex * t.wy, t.wx * t.wy};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int tl = off[k] - t0;
            if (!ok[k] || wt[k] == 0.f || (unsigned)tl >= (unsigned)PPB) continue;
            const int slot = atomicAdd(&s_cnt[tl], 1);
            if (slot < kSlotsS) {
                s_src[slot * PPB + tl] = src;
                s_w[slot * PPB + tl] = wt[k];
            } else {
                const int j = atomicAdd(&s_cnt[PPB], 1);
                o_tl[j] = tl;
                o_src[j] = src;
                o_w[j] = wt[k];
            }
        }
    }
    __syncthreads();

    const int pl = tid % PPB, cs = tid / PPB;
    const int rem = t0 + pl;
    const bool live = rem < HW;
    const size_t step = (size_t)NS * HW;
    constexpr int UN = 4;
    // g_flow partial sums of this thread's channel slice (same arithmetic as warp_bwd_kernel)
    float gix = 0.f, giy = 0.f;
    Taps t;
    if (live && gflow != nullptr) {
        const int y = dW.div(rem), xo = rem - y * W;
        t = make_taps_k((float)y, (float)xo, __ldg(fl + rem), __ldg(fl + HW + rem), H, W, kx, ky, border, rcp);
        const T *gb = gout + ((size_t)b * C + cs) * HW + rem;
        const T *p0 = x + ((size_t)b * C + cs) * HW + t.o_nw;
        const T *p1 = p0 + W;
#pragma unroll UN
        for (int c = cs; c < C; c += NS, gb += step, p0 += step, p1 += step) {
            const float go = ldf(gb);
            const float vnw = t.k_nw ? ldf(p0) : 0.f, vne = t.k_ne ? ldf(p0 + 1) : 0.f;
            const float vsw = t.k_sw ? ldf(p1) : 0.f, vse = t.k_se ? ldf(p1 + 1) : 0.f;
            gix = fmaf(go, fmaf(t.ey, vne - vnw, t.wy * (vse - vsw)), gix);
            giy = fmaf(go, fmaf(t.ex, vsw - vnw, t.wx * (vse - vne)), giy);
        }
    }
    red[cs * PPB + pl] = gix;
    red[kBlock + cs * PPB + pl] = giy;
    // slice 0 sorts its target's list by source pixel (in place)
    int n = min(s_cnt[pl], kSlotsS);
    int src[kSlotsS];
    float w[kSlotsS];
    if (cs == 0) {
#pragma unroll
        for (int i = 0; i < kSlotsS; ++i) {
            src[i] = i < n ? s_src[i * PPB + pl] : 0x7fffffff;
            w[i] = i < n ? s_w[i * PPB + pl] : 0.f;
        }
        if (n <= 8) {
#define USF_CS(a, b) cswap(src[a], w[a], src[b], w[b])
            USF_CS(0, 1); USF_CS(2, 3); USF_CS(4, 5); USF_CS(6, 7);
            USF_CS(0, 2); USF_CS(1, 3); USF_CS(4, 6); USF_CS(5, 7);
            USF_CS(1, 2); USF_CS(5, 6); USF_CS(0, 4); USF_CS(3, 7);
            USF_CS(1, 5); USF_CS(2, 6);
            USF_CS(1, 4); USF_CS(3, 6);
            USF_CS(2, 4); USF_CS(3, 5);
            USF_CS(3, 4);
#undef USF_CS
        } else {
            // Batcher odd-even merge sort, 16 keys, 63 comparators
#define USF_CS(a, b) cswap(src[a], w[a], src[b], w[b])
            USF_CS(0, 1); USF_CS(2, 3); USF_CS(4, 5); USF_CS(6, 7); USF_CS(8, 9); USF_CS(10, 11);
            USF_CS(12, 13); USF_CS(14, 15); USF_CS(0, 2); USF_CS(1, 3); USF_CS(4, 6); USF_CS(5, 7);
            USF_CS(8, 10); USF_CS(9, 11); USF_CS(12, 14); USF_CS(13, 15); USF_CS(1, 2); USF_CS(5, 6);
            USF_CS(9, 10); USF_CS(13, 14); USF_CS(0, 4); USF_CS(1, 5); USF_CS(2, 6); USF_CS(3, 7);
            USF_CS(8, 12); USF_CS(9, 13); USF_CS(10, 14); USF_CS(11, 15); USF_CS(2, 4); USF_CS(3, 5);
            USF_CS(10, 12); USF_CS(11, 13); USF_CS(1, 2); USF_CS(3, 4); USF_CS(5, 6); USF_CS(9, 10);
            USF_CS(11, 12); USF_CS(13, 14); USF_CS(0, 8); USF_CS(1, 9); USF_CS(2, 10); USF_CS(3, 11);
            USF_CS(4, 12); USF_CS(5, 13); USF_CS(6, 14); USF_CS(7, 15); USF_CS(4, 8); USF_CS(5, 9);
            USF_CS(6, 10); USF_CS(7, 11); USF_CS(2, 4); USF_CS(3, 5); USF_CS(6, 8); USF_CS(7, 9);
            USF_CS(10, 12); USF_CS(11, 13); USF_CS(1, 2); USF_CS(3, 4); USF_CS(5, 6); USF_CS(7, 8);
            USF_CS(9, 10); USF_CS(11, 12); USF_CS(13, 14);
#undef USF_CS
        }
        if (NS > 1) {
#pragma unroll
            for (int i = 0; i < kSlotsS; ++i) {
                s_src[i * PPB + pl] = src[i];
                s_w[i * PPB + pl] = w[i];
            }
        }
    }
    __syncthreads();
    if (!live) return;
    if (gflow != nullptr && cs == 0) {
        gix = red[pl];
        giy = red[kBlock + pl];
#pragma unroll
        for (int s2 = 1; s2 < NS; ++s2) {
            gix += red[s2 * PPB + pl];
            giy += red[kBlock + s2 * PPB + pl];
        }
        const float gxn = __fmul_rn(t.mx, gix), gyn = __fmul_rn(t.my, giy);
        const float dfx = rcp ? __fmul_rn(gxn, kx.rcp) : div_by_size(gxn, kx);
        const float dfy = rcp ? __fmul_rn(gyn, ky.rcp) : div_by_size(gyn, ky);
        gflow[(size_t)b * 2 * HW + rem] = __fmul_rn(dfx, 2.0f);
        gflow[(size_t)b * 2 * HW + HW + rem] = __fmul_rn(dfy, 2.0f);
    }
    if (cs != 0) {
#pragma unroll
        for (int i = 0; i < kSlotsS; ++i) {
            src[i] = s_src[i * PPB + pl];
            w[i] = s_w[i * PPB + pl];
        }
    }
    const int novf = s_cnt[PPB];
    const T *plane = gout + ((size_t)b * C + cs) * HW;
    T *o = gx + ((size_t)b * C + cs) * HW + rem;
#pragma unroll UN
    for (int c = cs; c < C; c += NS, plane += step, o += step) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < n) acc = fmaf(w[i], ldf(plane + src[i]), acc);
        if (n > 4) {
#pragma unroll
            for (int i = 4; i < 8; ++i)
                if (i < n) acc = fmaf(w[i], ldf(plane + src[i]), acc);
            if (n > 8) {
#pragma unroll
                for (int i = 8; i < kSlotsS; ++i)
                    if (i < n) acc = fmaf(w[i], ldf(plane + src[i]), acc);
            }
        }
        for (int j = 0; j < novf; ++j)  // rare: more than 16 taps on one target (flow collapsing onto a border)
            if (o_tl[j] == pl) acc = fmaf(o_w[j], ldf(plane + o_src[j]), acc);
        stf(o, acc);
    }
}

// Overflow side list: normally short, but NOT normally empty -- a flow that leaves the image is clipped onto the
// border, and a border pixel easily collects more than kSlots taps (the bench's 3-px flows put ~17 k entries here at
// 8x64x208).  Parallel over entries x channels; folding this pass into the gather's last CTA was tried and cost
// 150 us at that size.  The last CTA to finish resets the counters (USF_FLAG_WORKSPACE_CLEAN contract).
template <typename T>
__global__ void __launch_bounds__(kBlock)
warp_gx_overflow(const T *__restrict__ gout, T *__restrict__ gx, Workspace ws, int cap, int C, int HW) {
    pdl_wait();
    __shared__ int s_last;
    const int n = min(ws.ovf_n[0], cap);
    const long total = (long)n * C;
    for (long it = (long)blockIdx.x * kBlock + threadIdx.x; it < total; it += (long)gridDim.x * kBlock) {
        const int c = (int)(it / n), j = (int)(it - (long)c * n);
        const int tgt = ws.ovf_tgt[j];
        const int b = tgt / HW, trem = tgt - b * HW;
        const size_t plane = ((size_t)b * C + c) * HW;
        atomic_addf(gx + plane + trem, ws.ovf_w[j] * ldf(gout + plane + ws.ovf_src[j]));
    }
    __syncthreads();  // every thread of this CTA has read ovf_n[0]
    if (threadIdx.x == 0) s_last = atomicAdd(ws.ovf_n + 1, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        ws.ovf_n[0] = 0;
        ws.ovf_n[1] = 0;
    }
}

// Channel slices per pixel: enough threads to fill the chip (~one full wave of 148 x 2048), but as few
// slices as possible because every slice repeats the coordinate arithmetic (two IEEE divisions).  Small
// levels end up with 4-8 channels per thread: their cost is the depth of the dependent-load chain.
// images of at most this many pixels take the fused single-kernel backward (USF_WARP_SMALL_MAX_HW=0 disables)
const int kWarpSmallMaxHW = [] {
    const char *e = getenv("USF_WARP_SMALL_MAX_HW");
    const int v = e ? atoi(e) : 1024;
    return v > 1024 ? 1024 : v;
}();

int pick_ns(int C, int N) {
    // USF_WARP_NS=<power of two>: force the number of channel slices (tuning runs)
    static const int forced = [] { const char *e = getenv("USF_WARP_NS"); return e ? atoi(e) : 0; }();
    if (forced >= 1 && forced <= 32 && (forced & (forced - 1)) == 0 && C >= forced) return forced;
    // at least two slices once there are 16 channels: one thread walking 32 channels of a level that does not
    // fit in L2 (64x32x64x208) ran the gather at 10 % of DRAM bandwidth (ncu), two slices at twice that
    int ns = C >= 16 ? 2 : 1;
    const long wave = (long)sm_count() * 2048;
    while (ns < 32 && C >= 8 * ns && (ns < 8 ? (long)N * ns < wave : (long)N * ns * 2 <= wave)) ns *= 2;
    return ns;
}

}  // namespace

// images handled by the fused kernel need no workspace
size_t warp_bwd_workspace_bytes(int B, int H, int W) {
    return (long)H * W <= kWarpSmallMaxHW ? 0 : ws_bytes((size_t)B * H * W);
}
// the tap counters at the head of the workspace (what the per-call memset covers)
size_t warp_bwd_workspace_clean_bytes(int B, int H, int W) {
    return (long)H * W <= kWarpSmallMaxHW ? 0 : cnt_bytes((size_t)B * H * W) + 16;
}

// CTAs of `kernel` that fit on the device at once (cached per kernel and device)
static int resident_ctas(const void *kernel) {
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, int> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return sm_count() * 4;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(kernel, dev);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    return cache[key] = per_sm * sm_count();
}

template <typename T>
static int warp_fwd_t(const T *x, const float *flow, T *out, int B, int C, int H, int W, int pad_mode, unsigned flags,
                      cudaStream_t stream) {
    const int N = B * H * W;
    const int border = pad_mode == USF_PAD_BORDER, rcp = (flags & USF_FLAG_COORD_RECIPROCAL) ? 1 : 0;
    const int ns = pick_ns(C, N);
    const int blocks = (N + kBlock / ns - 1) / (kBlock / ns);
#define USF_LAUNCH(NS)                                                                      \
    {                                                                                       \
        const int res = resident_ctas((const void *)warp_fwd_kernel<NS, T>);                \
        usf_launch((warp_fwd_kernel<NS, T>), dim3(blocks < res ? blocks : res), dim3(kBlock), 0, stream, x, flow, out, N, C, H, W, border, rcp, make_fastdiv(H * W), make_fastdiv(W), make_axis(W), make_axis(H)); \
    }
    if (ns == 1 && C == 3) {  // loss-side warp of an RGB image
        const int res = resident_ctas((const void *)warp_rgb_kernel<T, false>);
        usf_launch((warp_rgb_kernel<T, false>), dim3(blocks < res ? blocks : res), dim3(kBlock), 0, stream, x, flow,
                   (const T *)nullptr, out, (float *)nullptr, N, H, W, border, rcp, make_fastdiv(H * W), make_fastdiv(W),
                   make_axis(W), make_axis(H));
        return note_launch(cudaGetLastError());
    }
    switch (ns) {
        case 32: USF_LAUNCH(32); break;
        case 16: USF_LAUNCH(16); break;
        case 8: USF_LAUNCH(8); break;
        case 4: USF_LAUNCH(4); break;
        case 2: USF_LAUNCH(2); break;
        default: USF_LAUNCH(1); break;
    }
#undef USF_LAUNCH
    return note_launch(cudaGetLastError());
}

template <int MODE, typename T>
static cudaError_t launch_bwd(int ns, int blocks, const T *x, const float *flow, const T *gout, T *gx, float *gflow,
                              Workspace ws, int N, int C, int H, int W, int border, int rcp, cudaStream_t stream) {
#define USF_LAUNCH(NS)                                                                                          \
    {                                                                                                           \
        const int res = resident_ctas((const void *)warp_bwd_kernel<NS, MODE, T>);                              \
        usf_launch((warp_bwd_kernel<NS, MODE, T>), dim3(blocks < res ? blocks : res), dim3(kBlock), 0, stream,                        \
            x, flow, gout, gx, gflow, ws, N, C, H, W, border, rcp, make_fastdiv(H * W), make_fastdiv(W), make_axis(W), make_axis(H)); \
    }
    switch (ns) {
        case 32: USF_LAUNCH(32); break;
        case 16: USF_LAUNCH(16); break;
        case 8: USF_LAUNCH(8); break;
        case 4: USF_LAUNCH(4); break;
        case 2: USF_LAUNCH(2); break;
        default: USF_LAUNCH(1); break;
    }
#undef USF_LAUNCH
    return cudaGetLastError();
}

template <typename T>
static int warp_bwd_t(const T *x, const float *flow, const T *gout, T *gx, float *gflow, int B, int C, int H, int W,
                      int pad_mode, unsigned flags, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
    const int N = B * H * W, HW = H * W;
    const int border = pad_mode == USF_PAD_BORDER, rcp = (flags & USF_FLAG_COORD_RECIPROCAL) ? 1 : 0;
    const int ns = pick_ns(C, N);
    const int blocks = (N + kBlock / ns - 1) / (kBlock / ns);
    Workspace ws = {};
    if (gx == nullptr) {
        if (ns == 1 && C == 3) {  // loss-side warp of an RGB image: gradient wrt the flow only
            const int res = resident_ctas((const void *)warp_rgb_kernel<T, true>);
            usf_launch((warp_rgb_kernel<T, true>), dim3(blocks < res ? blocks : res), dim3(kBlock), 0, stream, x, flow, gout,
                       (T *)nullptr, gflow, N, H, W, border, rcp, make_fastdiv(H * W), make_fastdiv(W), make_axis(W),
                       make_axis(H));
            return note_launch(cudaGetLastError());
        }
        return note_launch(launch_bwd<0, T>(ns, blocks, x, flow, gout, gx, gflow, ws, N, C, H, W, border, rcp, stream));
    }
    if (flags & USF_FLAG_WARP_BWD_ATOMIC) {
        cudaError_t e = cudaMemsetAsync(gx, 0, (size_t)N * C * sizeof(T), stream);
        if (e != cudaSuccess) { note_cuda_error(e); return USF_ERR_CUDA; }
        return note_launch(launch_bwd<2, T>(ns, blocks, x, flow, gout, gx, gflow, ws, N, C, H, W, border, rcp, stream));
    }
    if (HW <= kWarpSmallMaxHW) {  // one fused kernel, no workspace
        const int ppb = kBlock / ns, bpi = (HW + ppb - 1) / ppb;
        const size_t smem = ((size_t)(ppb + 4) + 2 * kSlotsS * ppb + 2 * kBlock + 3 * 4 * HW) * 4;
        cudaError_t e = cudaSuccess;
#define USF_LAUNCH(NS)                                                                                       \
    if ((e = ensure_smem((const void *)warp_bwd_small<NS, T>, (int)smem)) == cudaSuccess)                    \
        usf_launch((warp_bwd_small<NS, T>), dim3(B * bpi), dim3(kBlock), smem, stream, x, flow, gout, gx, gflow, C, H, W, border, rcp, \
                                                                 make_fastdiv(W), make_fastdiv(bpi), make_axis(W), make_axis(H));
        switch (ns) {
            case 32: USF_LAUNCH(32); break;
            case 16: USF_LAUNCH(16); break;
            case 8: USF_LAUNCH(8); break;
            case 4: USF_LAUNCH(4); break;
            case 2: USF_LAUNCH(2); break;
            default: USF_LAUNCH(1); break;
        }
#undef USF_LAUNCH
        return note_launch(e != cudaSuccess ? e : cudaGetLastError());
    }
    if (workspace == nullptr || workspace_bytes < ws_bytes((size_t)N) || ((uintptr_t)workspace & 15))
        return USF_ERR_WORKSPACE;
    ws = carve(workspace, (size_t)N);
    if (!(flags & USF_FLAG_WORKSPACE_CLEAN)) {  // counters must start at zero; the gather zeroes them again
        cudaError_t e = cudaMemsetAsync(ws.cnt, 0, cnt_bytes((size_t)N) + 16, stream);
        if (e != cudaSuccess) { note_cuda_error(e); return USF_ERR_CUDA; }
    }
    int st = note_launch(launch_bwd<1, T>(ns, blocks, x, flow, gout, gx, gflow, ws, N, C, H, W, border, rcp, stream));
    if (st != USF_OK) return st;
#define USF_LAUNCH(NS) usf_launch((warp_gx_gather<NS, T>), dim3(blocks), dim3(kBlock), 0, stream, gout, gx, ws, N, C, HW, make_fastdiv(HW))
    switch (ns) {
        case 32: USF_LAUNCH(32); break;
        case 16: USF_LAUNCH(16); break;
        case 8: USF_LAUNCH(8); break;
        case 4: USF_LAUNCH(4); break;
        case 2: USF_LAUNCH(2); break;
        default: USF_LAUNCH(1); break;
    }
#undef USF_LAUNCH
    st = note_launch(cudaGetLastError());
    if (st != USF_OK) return st;
    // fixed grid (the entry count lives on the device), sized for the worst case of this shape: up to 2 CTAs
    // per SM (the list is normally short), grid-stride over entries x channels
    long ovf_blocks = ((long)4 * N * C + 16L * kBlock - 1) / (16L * kBlock);
    ovf_blocks = ovf_blocks < 16 ? 16 : (ovf_blocks > sm_count() * 2 ? sm_count() * 2 : ovf_blocks);
    usf_launch(warp_gx_overflow<T>, dim3((int)ovf_blocks), dim3(kBlock), 0, stream, gout, gx, ws, 4 * N, C, HW);
    return note_launch(cudaGetLastError());
}

int warp_fwd_f32(const float *x, const float *flow, float *out, int B, int C, int H, int W, int pad_mode,
                 unsigned flags, cudaStream_t stream) {
    return warp_fwd_t<float>(x, flow, out, B, C, H, W, pad_mode, flags, stream);
}
int warp_bwd_f32(const float *x, const float *flow, const float *gout, float *gx, float *gflow, int B, int C,
                 int H, int W, int pad_mode, unsigned flags, void *workspace, size_t workspace_bytes,
                 cudaStream_t stream) {
    return warp_bwd_t<float>(x, flow, gout, gx, gflow, B, C, H, W, pad_mode, flags, workspace, workspace_bytes, stream);
}
// bf16 storage of x / out / gout / gx; flow, coordinates, weights and all sums stay fp32
int warp_fwd_bf16(const void *x, const float *flow, void *out, int B, int C, int H, int W, int pad_mode,
                  unsigned flags, cudaStream_t stream) {
    return warp_fwd_t<bf16>((const bf16 *)x, flow, (bf16 *)out, B, C, H, W, pad_mode, flags, stream);
}
int warp_bwd_bf16(const void *x, const float *flow, const void *gout, void *gx, float *gflow, int B, int C, int H,
                  int W, int pad_mode, unsigned flags, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
    return warp_bwd_t<bf16>((const bf16 *)x, flow, (const bf16 *)gout, (bf16 *)gx, gflow, B, C, H, W, pad_mode, flags,
                            workspace, workspace_bytes, stream);
}

}  // namespace usf
