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

// Launch-time constants of one image axis (size-1 and friends), passed to the kernels by value.
struct AxisK {
    float sz;    // size - 1
    float rcp;   // RN(1 / sz)
    float half;  // sz * 0.5 (exact): d(ix)/d(normalised coordinate)
    int fast;    // sz <= 16400: t / sz by one FMA correction step, which is then EXACTLY the IEEE quotient
};
inline AxisK make_axis(int size) {
    AxisK k;
    k.sz = (float)(size - 1);
    k.rcp = 1.0f / k.sz;  // host IEEE division: correctly rounded
    k.half = k.sz * 0.5f;
    k.fast = (size - 1) <= 16400 ? 1 : 0;
    return k;
}

// Correctly rounded t / sz without the ~12-instruction IEEE division sequence.  With y = RN(1/sz):
// q0 = RN(t*y), r = t - sz*q0 (exact in one FMA), q1 = RN(q0 + r*y).  For every integer divisor 1..16400 and
// EVERY fp32 significand of t this equals RN(t/sz) (exhaustive check: benchmarks/exact_div_check.c, run by
// tests/test_exact_division.py) -- the two IEEE divisions were a quarter of the loss-side warp's instructions.
// The quotient is scale invariant, so only t's exponent range matters: |t| >= 2^100, inf and NaN take the IEEE
// sequence (r would be inf - inf); tiny |t| needs no care because the caller subtracts 1 from the quotient.
__device__ __forceinline__ float div_by_size(float t, const AxisK &k) {
    if (k.fast && fabsf(t) < 1.2676506e30f) {
        const float q0 = __fmul_rn(t, k.rcp);
        const float r0 = __fmaf_rn(-k.sz, q0, t);
        return __fmaf_rn(r0, k.rcp, q0);
    }
    return __fdiv_rn(t, k.sz);
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

// The same pipeline with the launch-time constants precomputed and the quotient from div_by_size: bit-identical
// to src_coord (both are the correctly rounded IEEE result), about a third of the instructions.
__device__ __forceinline__ float src_coord_k(float fi, float fl, const AxisK &k, bool border, bool rcp, float &mult) {
    const float v = __fadd_rn(fi, fl);
    const float t = __fmul_rn(2.0f, v);
    const float qn = rcp ? __fmul_rn(t, k.rcp) : div_by_size(t, k);
    const float g = __fsub_rn(qn, 1.0f);
    const float h = __fmul_rn(__fadd_rn(g, 1.0f), 0.5f);
    float ix = __fmul_rn(h, k.sz);
    mult = k.half;
    if (border) {
        if (ix <= 0.f) { ix = 0.f; mult = 0.f; }
        else if (ix >= k.sz) { ix = k.sz; mult = 0.f; }
    }
    return ix;
}

__device__ __forceinline__ Taps taps_from_coords(float ix, float iy, int H, int W);

__device__ __forceinline__ Taps make_taps_k(float fy, float fx, float u, float v, int H, int W, const AxisK &kx,
                                            const AxisK &ky, bool border, bool rcp) {
    float mx, my;
    const float ix = src_coord_k(fx, u, kx, border, rcp, mx);
    const float iy = src_coord_k(fy, v, ky, border, rcp, my);
    Taps t = taps_from_coords(ix, iy, H, W);
    t.mx = mx;
    t.my = my;
    return t;
}

__device__ __forceinline__ Taps make_taps(int y, int xo, float u, float v, int H, int W, bool border,
                                          bool rcp) {
    float mx, my;
    const float ix = src_coord(xo, u, W, border, rcp, mx);
    const float iy = src_coord(y, v, H, border, rcp, my);
    Taps t = taps_from_coords(ix, iy, H, W);
    t.mx = mx;
    t.my = my;
    return t;
}

__device__ __forceinline__ Taps taps_from_coords(float ix, float iy, int H, int W) {
    Taps t;
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
    const bool xa = (unsigned)x0 < (unsigned)W, xb = (unsigned)x1 < (unsigned)W;
    const bool ya = (unsigned)y0 < (unsigned)H, yb = (unsigned)y1 < (unsigned)H;
    const bool fin = (ix == ix) && (iy == iy);  // NaN coordinates sample nothing
    t.k_nw = xa && ya && fin; t.k_ne = xb && ya && fin; t.k_sw = xa && yb && fin; t.k_se = xb && yb && fin;
    t.o_nw = y0 * W + x0; t.o_ne = t.o_nw + 1; t.o_sw = t.o_nw + W; t.o_se = t.o_sw + 1;
    return t;
}

}  // namespace usf
