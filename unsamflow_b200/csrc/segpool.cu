// segpool.cu -- mask-feature pooling of the "+mf" configurations.
//
// Replaces models/pwclite.py:317-340 (both images):
//     oh      = one_hot(seg)                                   [B,1,h,w,N]
//     pooled  = amax(oh * feat[..., None], dim=(2, 3))         [B,C,N]      (temporaries [B,C,h,w,N]: gigabytes)
//     out     = (oh * pooled[:, :, None, None, :]).sum(-1)     [B,C,h,w]
// i.e. every pixel receives the maximum of `feat` over its own segment.  In the product the pixels OUTSIDE a segment
// contribute zeros to the amax, so the pooled value is max(0, max over the segment) -- unless the segment covers the
// whole image, when there is no zero to compete with.  Reproduced exactly: tables start at 0 (at -inf for an
// all-covering segment).
//
// Forward (segpool_fwd_kernel): CTA = (image, group of kCh channels); the tables live in shared memory as
// order-preserving integer keys, filled with 32-bit shared-memory atomicMax (max is order independent: the result is
// bit-reproducible), the arg-max is the SMALLEST pixel index attaining the maximum (atomicMin in a second sweep).
// Backward (segpool_bwd_kernel): g_feat[c, p] = sum of g_out[c, .] over p's segment if p is that segment's arg-max,
// else 0.  The per-segment sums are order DEPENDENT in floating point, so they are formed without atomics: a warp
// sweeps its own contiguous range of pixels 32 at a time, lanes that share a segment are summed in lane order (shuffles),
// the group leader adds to the warp's private table, and the tables of the 8 warps are added in warp order.
// Deviation from autograd of the one-hot formulation, on a measure-zero set: when several pixels tie for the maximum
// (or an in-segment value equals the competing 0 exactly) torch.amax splits the gradient evenly among them; here the
// first pixel takes it (the clamp takes it -- no gradient -- when the 0 wins or ties).
#include "common.cuh"

namespace usf {
namespace {

constexpr int kSpBlock = 256;
constexpr int kSpChMax = 4;    // channels per CTA: 4, 2 or 1 -- fewer when that is what it takes to cover the SMs
constexpr int kSpWarps = kSpBlock / 32;

__device__ __forceinline__ int float_key(float f) {  // monotonic: a < b  <=>  key(a) < key(b)   (NaN sorts above +inf)
    const int b = __float_as_int(f);
    return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

// smem: int cnt[N]; int key[kSpCh][N]; int arg[kSpCh][N]
// Two sweeps over the image.  Sweep 1: pixel count and in-segment maximum per segment.  Sweep 2: arg-max and the
// un-pooled output.  Segment maps are piecewise constant, so the 32 pixels of a warp mostly share one segment and
// would serialise on one shared-memory word: lanes are grouped by segment (match.any) and reduced in registers
// (redux.sync) first; one lane per group touches shared memory.
template <typename T, int kSpCh>
__global__ void __launch_bounds__(kSpBlock)
segpool_fwd_kernel(const T *__restrict__ feat, const int *__restrict__ seg, T *__restrict__ out,
                   int *__restrict__ argmax, int C, int HW, int N) {
    pdl_wait();
    extern __shared__ int sp_smem[];
    int *cnt = sp_smem, *key = cnt + N, *arg = key + kSpCh * N;
    const int b = blockIdx.x, c0 = blockIdx.y * kSpCh, tid = threadIdx.x, lane = tid & 31;
    const int nc = min(kSpCh, C - c0);
    const int *sg = seg + (size_t)b * HW;
    const T *f = feat + ((size_t)b * C + c0) * HW;
    const int kneg = float_key(-INFINITY);
    for (int i = tid; i < N; i += kSpBlock) cnt[i] = 0;
    for (int i = tid; i < kSpCh * N; i += kSpBlock) {
        key[i] = kneg;
        arg[i] = 0x7fffffff;
    }
    __syncthreads();
    // whole warps (the votes need every lane), four iterations' loads in flight at once: the sweeps are bound by the
    // latency of one dependent load per iteration otherwise
    constexpr int U = 4;
    const int sweep = (HW + U * kSpBlock - 1) / (U * kSpBlock) * (U * kSpBlock);
    for (int p0 = tid; p0 < sweep; p0 += U * kSpBlock) {
        int sv[U], kv[U][kSpCh];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = p0 + u * kSpBlock;
            // (ids outside [0, N) are not segments: such a pixel joins no table and receives 0 / passes no gradient)
            int s = p < HW ? sg[p] : -1;
            if ((unsigned)s >= (unsigned)N) s = -1;
            sv[u] = s;
#pragma unroll
            for (int k = 0; k < kSpCh; ++k)
                kv[u][k] = (p < HW && k < nc) ? float_key(ldf(f + (size_t)k * HW + p)) : kneg;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int s = sv[u];
            // (a redux.sync over a PARTIAL mask is emulated by a loop: only the all-lanes-equal case, one instruction,
            // is worth taking; mixed warps fall back to one atomic per lane)
            if (__all_sync(0xffffffffu, s == __shfl_sync(0xffffffffu, s, 0))) {
                if (s < 0) continue;  // uniform
                if (lane == 0) atomicAdd(&cnt[s], 32);
#pragma unroll
                for (int k = 0; k < kSpCh; ++k) {
                    const int best = __reduce_max_sync(0xffffffffu, kv[u][k]);
                    if (lane == 0 && k < nc) atomicMax(&key[k * N + s], best);
                }
            } else if (s >= 0) {
                atomicAdd(&cnt[s], 1);
#pragma unroll
                for (int k = 0; k < kSpCh; ++k)
                    if (k < nc) atomicMax(&key[k * N + s], kv[u][k]);
            }
        }
    }
    __syncthreads();
    T *o = out + ((size_t)b * C + c0) * HW;
    for (int p0 = tid; p0 < sweep; p0 += U * kSpBlock) {
        int sv[U], kv[U][kSpCh];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = p0 + u * kSpBlock;
            int s = p < HW ? sg[p] : -1;
            if ((unsigned)s >= (unsigned)N) s = -1;
            sv[u] = s;
#pragma unroll
            for (int k = 0; k < kSpCh; ++k)
                kv[u][k] = (argmax != nullptr && p < HW && k < nc) ? float_key(ldf(f + (size_t)k * HW + p)) : kneg;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = p0 + u * kSpBlock, s = sv[u];
            const bool uni = __all_sync(0xffffffffu, s == __shfl_sync(0xffffffffu, s, 0));
            const bool whole = s >= 0 && cnt[s] == HW;  // the segment covers the image: no zero from outside competes
#pragma unroll
            for (int k = 0; k < kSpCh; ++k) {
                const int top = (s >= 0 && k < nc) ? key[k * N + s] : 0;
                if (argmax != nullptr) {
                    const int mine = (s >= 0 && k < nc && kv[u][k] == top) ? p : 0x7fffffff;
                    if (uni) {
                        const int first = __reduce_min_sync(0xffffffffu, mine);
                        if (lane == 0 && first != 0x7fffffff) atomicMin(&arg[k * N + s], first);
                    } else if (mine != 0x7fffffff) {
                        atomicMin(&arg[k * N + s], mine);
                    }
                }
                // key(0.0f) == 0: the clamp at zero is a max of integer keys
                if (p < HW && k < nc) stf(o + (size_t)k * HW + p, s >= 0 ? key_float(whole ? top : max(top, 0)) : 0.f);
            }
        }
    }
    if (argmax == nullptr) return;
    __syncthreads();
    // -1: the clamp at zero won or tied (no pixel receives the gradient), or the class is empty
    for (int i = tid; i < nc * N; i += kSpBlock) {
        const int n = i % N;
        const bool taken = cnt[n] > 0 && (cnt[n] == HW || key[i] > 0);
        argmax[((size_t)b * C + c0) * N + i] = taken ? arg[i] : -1;
    }
}

// smem: float part[kSpWarps][kSpCh][N]
template <typename T, int kSpCh>
__global__ void __launch_bounds__(kSpBlock)
segpool_bwd_kernel(const T *__restrict__ gout, const int *__restrict__ seg, const int *__restrict__ argmax,
                   T *__restrict__ gfeat, int C, int HW, int N) {
    pdl_wait();
    extern __shared__ float spb_smem[];
    const int b = blockIdx.x, c0 = blockIdx.y * kSpCh, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nc = min(kSpCh, C - c0);
    const int *sg = seg + (size_t)b * HW;
    const T *g = gout + ((size_t)b * C + c0) * HW;
    float *mine = spb_smem + (size_t)wid * kSpCh * N;
    for (int i = tid; i < kSpWarps * kSpCh * N; i += kSpBlock) spb_smem[i] = 0.f;
    __syncthreads();
    // warp `wid` owns pixels [lo, hi): a fixed partition, swept in increasing order
    const int per = (HW + kSpWarps - 1) / kSpWarps, lo = wid * per, hi = min(HW, lo + per);
    constexpr int U = 4;  // four groups of 32 pixels loaded together, then reduced one after the other (in order)
    for (int base0 = lo; base0 < hi; base0 += 32 * U) {
        int sv[U];
        float vv[U][kSpCh];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = base0 + 32 * u + lane;
            int s = p < hi ? sg[p] : -1;
            if ((unsigned)s >= (unsigned)N) s = -1;
            sv[u] = s;
#pragma unroll
            for (int k = 0; k < kSpCh; ++k) vv[u][k] = (s >= 0 && k < nc) ? ldf(g + (size_t)k * HW + p) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (base0 + 32 * u >= hi) break;  // warp-uniform
            const int s = sv[u];
            const bool on = s >= 0;
            const unsigned grp = __match_any_sync(0xffffffffu, s);
            const int leader = __ffs(grp) - 1;
            // sum the lanes of each group in lane order: every lane walks its own group's members (the lanes of a
            // group walk the same list, so they compute the same sum); warps are usually inside ONE segment
            float sum[kSpCh];
#pragma unroll
            for (int k = 0; k < kSpCh; ++k) sum[k] = 0.f;
            if (grp == 0xffffffffu) {
                // all 32 pixels in one segment (every lane sees the full mask, so the branch is uniform): a butterfly,
                // whose order is fixed too
#pragma unroll
                for (int k = 0; k < kSpCh; ++k) {
                    float t = vv[u][k];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                    sum[k] = t;
                }
            } else {
                // uniform trip count: the size of the largest group in the warp
                int most = __popc(grp);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) most = max(most, __shfl_xor_sync(0xffffffffu, most, o));
                unsigned rest = grp;
                for (int it = 0; it < most; ++it) {
                    const int srcl = rest ? __ffs(rest) - 1 : lane;
#pragma unroll
                    for (int k = 0; k < kSpCh; ++k) {
                        const float t = __shfl_sync(0xffffffffu, vv[u][k], srcl);
                        if (rest) sum[k] += t;
                    }
                    rest &= rest - 1;
                }
            }
            if (on && lane == leader) {
#pragma unroll
                for (int k = 0; k < kSpCh; ++k) mine[k * N + s] += sum[k];  // single writer per (warp, segment) at a time
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // tables of the 8 warps added in warp order, into warp 0's table
    for (int i = tid; i < kSpCh * N; i += kSpBlock) {
        float t = spb_smem[i];
#pragma unroll
        for (int w = 1; w < kSpWarps; ++w) t += spb_smem[(size_t)w * kSpCh * N + i];
        spb_smem[i] = t;
    }
    __syncthreads();
    const int *am = argmax + ((size_t)b * C + c0) * N;
    T *o = gfeat + ((size_t)b * C + c0) * HW;
#pragma unroll 4
    for (int p = tid; p < HW; p += kSpBlock) {
        const int s = sg[p];
        const bool in = (unsigned)s < (unsigned)N;
#pragma unroll
        for (int k = 0; k < kSpCh; ++k)
            if (k < nc) stf(o + (size_t)k * HW + p, (in && am[k * N + s] == p) ? spb_smem[k * N + s] : 0.f);
    }
}

inline int pick_channels(int B, int C) {
    int ch = kSpChMax;
    while (ch > 1 && (long)B * ((C + ch - 1) / ch) < 2L * sm_count()) ch /= 2;
    return ch;
}

template <typename T, int CH>
int segpool_fwd_c(const T *feat, const int *seg, T *out, int *argmax, int B, int C, int HW, int N, cudaStream_t stream) {
    const size_t smem = (size_t)(1 + 2 * CH) * N * sizeof(int);
    if (cudaError_t e = ensure_smem((const void *)segpool_fwd_kernel<T, CH>, (int)smem)) return note_launch(e);
    usf_launch((segpool_fwd_kernel<T, CH>), dim3(B, (C + CH - 1) / CH), dim3(kSpBlock), smem, stream, feat, seg, out, argmax,
               C, HW, N);
    return note_launch(cudaGetLastError());
}
template <typename T>
int segpool_fwd_t(const T *feat, const int *seg, T *out, int *argmax, int B, int C, int HW, int N, cudaStream_t stream) {
    switch (pick_channels(B, C)) {
        case 4: return segpool_fwd_c<T, 4>(feat, seg, out, argmax, B, C, HW, N, stream);
        case 2: return segpool_fwd_c<T, 2>(feat, seg, out, argmax, B, C, HW, N, stream);
        default: return segpool_fwd_c<T, 1>(feat, seg, out, argmax, B, C, HW, N, stream);
    }
}
template <typename T, int CH>
int segpool_bwd_c(const T *gout, const int *seg, const int *argmax, T *gfeat, int B, int C, int HW, int N,
                  cudaStream_t stream) {
    const size_t smem = (size_t)kSpWarps * CH * N * sizeof(float);
    if (cudaError_t e = ensure_smem((const void *)segpool_bwd_kernel<T, CH>, (int)smem)) return note_launch(e);
    usf_launch((segpool_bwd_kernel<T, CH>), dim3(B, (C + CH - 1) / CH), dim3(kSpBlock), smem, stream, gout, seg, argmax,
               gfeat, C, HW, N);
    return note_launch(cudaGetLastError());
}
template <typename T>
int segpool_bwd_t(const T *gout, const int *seg, const int *argmax, T *gfeat, int B, int C, int HW, int N,
                  cudaStream_t stream) {
    switch (pick_channels(B, C)) {
        case 4: return segpool_bwd_c<T, 4>(gout, seg, argmax, gfeat, B, C, HW, N, stream);
        case 2: return segpool_bwd_c<T, 2>(gout, seg, argmax, gfeat, B, C, HW, N, stream);
        default: return segpool_bwd_c<T, 1>(gout, seg, argmax, gfeat, B, C, HW, N, stream);
    }
}

}  // namespace

// largest segment count the shared-memory tables hold (backward: 8 warps x 4 channels x N floats <= 200 KB)
int segpool_max_segments() { return 1536; }

int segpool_fwd(const void *feat, const int *seg, void *out, int *argmax, int B, int C, int HW, int N, int dtype,
                cudaStream_t stream) {
    if (dtype == USF_F32) return segpool_fwd_t<float>((const float *)feat, seg, (float *)out, argmax, B, C, HW, N, stream);
    if (dtype == USF_BF16) return segpool_fwd_t<bf16>((const bf16 *)feat, seg, (bf16 *)out, argmax, B, C, HW, N, stream);
    return USF_ERR_BAD_ARG;
}
int segpool_bwd(const void *gout, const int *seg, const int *argmax, void *gfeat, int B, int C, int HW, int N, int dtype,
                cudaStream_t stream) {
    if (dtype == USF_F32) return segpool_bwd_t<float>((const float *)gout, seg, argmax, (float *)gfeat, B, C, HW, N, stream);
    if (dtype == USF_BF16) return segpool_bwd_t<bf16>((const bf16 *)gout, seg, argmax, (bf16 *)gfeat, B, C, HW, N, stream);
    return USF_ERR_BAD_ARG;
}

}  // namespace usf
