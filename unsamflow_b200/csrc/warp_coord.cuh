// warp_coord.cuh -- the coordinate pipeline of flow_warp, shared by warp.cu and occ.cu.
//
// utils/warp_utils.py:7-23,97-106 (mesh_grid + norm_grid + grid_sample's unnormalise / clip) as separately
// rounded fp32 operations in the reference's order, and the four bilinear taps that follow from it.
#pragma once
#include "common.cuh"

namespace usf {

// Division by a launch-time constant: q = floor(n / d) for 0 <= n < 2^31 as one IMAD.HI + add + shift
// (M = ceil(2^(32+L) / d) = 2^32 + m with L = ceil(log2 d); the integer divides were a quarter of all
// instructions of the 3-channel loss-side warp).
struct FastDiv {
    unsigned m;
    int L;
    int d;
    __device__ __forceinline__ int div(int n) const { return (int)((__umulhi(m, (unsigned)n) + (unsigned)n) >> L); }
};
inline FastDiv make_fastdiv(int d) {
    FastDiv f;
    f.d = d;
    f.L = 0;
    while ((1u << f.L) < (unsigned)d) ++f.L;
    const unsigned long long one = 1ull << (32 + f.L);
    const unsigned long long M = one / (unsigned)d + ((one % (unsigned)d) ? 1 : 0);
    f.m = (unsigned)(M - (1ull << 32));
    return f;
}

struct Taps {
    int o_nw, o_ne, o_sw, o_se;      // offsets inside one H*W plane (valid only if ok_*)
    bool k_nw, k_ne, k_sw, k_se;     // tap inside the image
    float ex, wx, ey, wy;            // 1-D weights: ex = x1 - ix, wx = ix - x0, ...
    float mx, my;                    // d(ix)/d(grid) multipliers incl. border clip gradient
};

// One axis of the reference's coordinate pipeline.  rcp: ATen-CUDA flavour of `/ (size-1)`.
__device__ __forceinline__ float src_coord(int i, float fl, int size, bool border, bool rcp, float &mult) {
    const float sz = (float)(size - 1);
    const float v = __fadd_rn((float)i, fl);
    const float t = __fmul_rn(2.0f, v);
    const float qn = rcp ? __fmul_rn(t, __fdiv_rn(1.0f, sz)) : __fdiv_rn(t, sz);
    const float g = __fsub_rn(qn, 1.0f);
    const float h = __fmul_rn(__fadd_rn(g, 1.0f), 0.5f);
    float ix = __fmul_rn(h, sz);
    mult = __fmul_rn(sz, 0.5f);
    if (border) {
        if (ix <= 0.f) { ix = 0.f; mult = 0.f; }
        else if (ix >= sz) { ix = sz; mult = 0.f; }
    }
    return ix;
}

__device__ __forceinline__ Taps make_taps(int y, int xo, float u, float v, int H, int W, bool border,
                                          bool rcp) {
    Taps t;
    const float ix = src_coord(xo, u, W, border, rcp, t.mx);
    const float iy = src_coord(y, v, H, border, rcp, t.my);
    // floorf on +-inf / NaN: the int conversions saturate and every tap is out of range
    const float fx = floorf(ix), fy = floorf(iy);
    t.ex = __fsub_rn(__fadd_rn(fx, 1.0f), ix);
    t.wx = __fsub_rn(ix, fx);
    t.ey = __fsub_rn(__fadd_rn(fy, 1.0f), iy);
    t.wy = __fsub_rn(iy, fy);
    // clamp before the int conversion so huge flows cannot overflow the offset arithmetic
    const int x0 = (int)fminf(fmaxf(fx, -2.0f), (float)W);
    const int y0 = (int)fminf(fmaxf(fy, -2.0f), (float)H);
    const int x1 = x0 + 1, y1 = y0 + 1;
    const bool xa = (x0 >= 0) && (x0 < W), xb = (x1 >= 0) && (x1 < W);
    const bool ya = (y0 >= 0) && (y0 < H), yb = (y1 >= 0) && (y1 < H);
    const bool fin = (ix == ix) && (iy == iy);  // NaN coordinates sample nothing
    t.k_nw = xa && ya && fin; t.k_ne = xb && ya && fin; t.k_sw = xa && yb && fin; t.k_se = xb && yb && fin;
    t.o_nw = y0 * W + x0; t.o_ne = t.o_nw + 1; t.o_sw = t.o_nw + W; t.o_se = t.o_sw + 1;
    return t;
}

}  // namespace usf
