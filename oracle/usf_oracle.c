/*
 * usf_oracle.c -- CPU restatement of UnSAMFlow's PWC-Lite hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path (unsamflow_b200/) never does.
 *
 * Parity pinning: the reference holds no golden vectors for this path
 * (SURVEY.md 8c).  This restatement is pinned against outputs of the reference
 * itself, generated in the build container by oracle/make_golden.py (imports
 * /root/reference/models/correlation_native.py and utils/warp_utils.py) and
 * committed as tests/golden/ *.npz.  tests/test_oracle_golden.py checks every
 * function below against them.
 *
 * What is restated (reference file:line, relative to the reference root):
 *   usf_oracle_corr_fwd   models/correlation_native.py:13-23  (definition)
 *                         cross-checked by models/correlation_package/
 *                         correlation_cuda_kernel.cu:82-109
 *   usf_oracle_corr_bwd   autograd of the above; same sums as
 *                         correlation_cuda_kernel.cu:180-205 (grad input1) and
 *                         :252-298 (grad input2)
 *   usf_oracle_warp_fwd   utils/warp_utils.py:7-13 (mesh_grid), :16-23
 *                         (norm_grid), :97-106 (flow_warp) + the ATen
 *                         grid_sampler_2d arithmetic it calls (third-party:
 *                         PyTorch, unpinned by the reference; helper math in
 *                         torch/include/ATen/native/cuda/GridSampler.cuh:21-31
 *                         unnormalize, :53-57 clip, :219-246 bounds)
 *   usf_oracle_warp_bwd   autograd of the above (GridSampler.cuh:62-80 clip
 *                         gradient) chained through norm_grid.
 *   usf_oracle_occ_bidirection / usf_oracle_occ_backward
 *                         utils/warp_utils.py:109-117 / :26-94,120-126 (the two
 *                         occlusion masks; see the block at the end of this file)
 *
 * Arithmetic: every coordinate operation is a separately rounded fp32 operation
 * in the reference's order (compile with -ffp-contract=off, no -ffast-math).
 * Channel sums accumulate in double and round once, so this oracle sits at the
 * centre of the fp32 summation-order noise of the reference (~1e-7 relative).
 *
 * Layout: NCHW contiguous fp32; flow is [B,2,H,W] with channel 0 = u (along W)
 * and channel 1 = v (along H).  pad_mode: 0 = "border", 1 = "zeros".
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define IDX4(b, c, y, x, C, H, W) ((((size_t)(b) * (C) + (c)) * (H) + (y)) * (size_t)(W) + (x))

int usf_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void usf_oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* out[b, (dy+md)*D + (dx+md), y, x] = (1/C) * sum_c x1[b,c,y,x] * x2[b,c,y+dy,x+dx]
 * with x2 read as zero outside the image (F.pad zeros, correlation_native.py:16). */
void usf_oracle_corr_fwd(const float *x1, const float *x2, float *out, int B, int C, int H, int W,
                         int md) {
    const int D = 2 * md + 1;
    const int K = D * D;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int y = 0; y < H; ++y) {
            for (int k = 0; k < K; ++k) {
                const int dy = k / D - md, dx = k % D - md;
                const int yy = y + dy;
                float *orow = out + IDX4(b, k, y, 0, K, H, W);
                for (int x = 0; x < W; ++x) {
                    const int xx = x + dx;
                    double acc = 0.0;
                    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                        for (int c = 0; c < C; ++c)
                            acc += (double)x1[IDX4(b, c, y, x, C, H, W)] *
                                   (double)x2[IDX4(b, c, yy, xx, C, H, W)];
                    }
                    orow[x] = (float)(acc / (double)C);
                }
            }
        }
    }
}

/* gx1[b,c,y,x] = (1/C) sum_k gout[b,k,y,x]     * x2[b,c,y+dy,x+dx]   (x2 zero outside)
 * gx2[b,c,y,x] = (1/C) sum_k gout[b,k,y-dy,x-dx] * x1[b,c,y-dy,x-dx]  (terms outside dropped)
 * Either output pointer may be NULL. */
void usf_oracle_corr_bwd(const float *x1, const float *x2, const float *gout, float *gx1,
                         float *gx2, int B, int C, int H, int W, int md) {
    const int D = 2 * md + 1;
    const int K = D * D;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int c = 0; c < C; ++c) {
            for (int y = 0; y < H; ++y) {
                for (int x = 0; x < W; ++x) {
                    double a1 = 0.0, a2 = 0.0;
                    for (int k = 0; k < K; ++k) {
                        const int dy = k / D - md, dx = k % D - md;
                        if (gx1) {
                            const int yy = y + dy, xx = x + dx;
                            if (yy >= 0 && yy < H && xx >= 0 && xx < W)
                                a1 += (double)gout[IDX4(b, k, y, x, K, H, W)] *
                                      (double)x2[IDX4(b, c, yy, xx, C, H, W)];
                        }
                        if (gx2) {
                            const int ys = y - dy, xs = x - dx;
                            if (ys >= 0 && ys < H && xs >= 0 && xs < W)
                                a2 += (double)gout[IDX4(b, k, ys, xs, K, H, W)] *
                                      (double)x1[IDX4(b, c, ys, xs, C, H, W)];
                        }
                    }
                    if (gx1) gx1[IDX4(b, c, y, x, C, H, W)] = (float)(a1 / (double)C);
                    if (gx2) gx2[IDX4(b, c, y, x, C, H, W)] = (float)(a2 / (double)C);
                }
            }
        }
    }
}

/* Source coordinate of output pixel (y, x) along one axis, replicating the
 * reference's fp32 op sequence:
 *   v  = float(i) + flow                      warp_utils.py:102  (base_grid + flow12)
 *   g  = 2.0*v / (size-1) - 1.0               warp_utils.py:21-22 (norm_grid)
 *   ix = ((g + 1) / 2) * (size-1)             GridSampler.cuh:21-31, align_corners=True
 * border: clip to [0, size-1] and zero the gradient multiplier at/after the
 * limits (GridSampler.cuh:53-57, :62-80).
 * *mult receives d(ix)/d(flow) as autograd would chain it:
 *   grad_flow = ((gix * mult) / (size-1)) * 2  with mult = (size-1)/2 * clip_grad. */
static inline float src_coord(int i, float fl, int size, int pad_mode, float *mult) {
    volatile float v = (float)i + fl;
    volatile float t = 2.0f * v;
    volatile float q = t / (float)(size - 1);
    volatile float g = q - 1.0f;
    volatile float h = g + 1.0f;
    volatile float h2 = h / 2.0f;
    volatile float ix = h2 * (float)(size - 1);
    float m = (float)(size - 1) / 2.0f;
    float r = ix;
    if (pad_mode == 0) {
        /* clip_coordinates_set_grad: in <= 0 -> (0, grad 0); in >= max -> (max, grad 0) */
        const float mx = (float)(size - 1);
        if (r <= 0.0f) {
            r = 0.0f;
            m = 0.0f;
        } else if (r >= mx) {
            r = mx;
            m = 0.0f;
        }
    }
    if (mult) *mult = m;
    return r;
}

static inline int inb(int y, int x, int H, int W) { return y >= 0 && y < H && x >= 0 && x < W; }

void usf_oracle_warp_fwd(const float *x, const float *flow, float *out, int B, int C, int H, int W,
                         int pad_mode) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int y = 0; y < H; ++y) {
            for (int xo = 0; xo < W; ++xo) {
                const float u = flow[IDX4(b, 0, y, xo, 2, H, W)];
                const float v = flow[IDX4(b, 1, y, xo, 2, H, W)];
                const float ix = src_coord(xo, u, W, pad_mode, NULL);
                const float iy = src_coord(y, v, H, pad_mode, NULL);
                const float fx = floorf(ix), fy = floorf(iy);
                const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
                /* GridSampler: nw = (ix_se - ix) * (iy_se - iy), ... */
                const float ex = (fx + 1.0f) - ix, wx = ix - fx;
                const float ey = (fy + 1.0f) - iy, wy = iy - fy;
                const float nw = ex * ey, ne = wx * ey, sw = ex * wy, se = wx * wy;
                for (int c = 0; c < C; ++c) {
                    const float *p = x + IDX4(b, c, 0, 0, C, H, W);
                    double acc = 0.0;
                    if (inb(y0, x0, H, W)) acc += (double)(p[(size_t)y0 * W + x0] * nw);
                    if (inb(y0, x1, H, W)) acc += (double)(p[(size_t)y0 * W + x1] * ne);
                    if (inb(y1, x0, H, W)) acc += (double)(p[(size_t)y1 * W + x0] * sw);
                    if (inb(y1, x1, H, W)) acc += (double)(p[(size_t)y1 * W + x1] * se);
                    out[IDX4(b, c, y, xo, C, H, W)] = (float)acc;
                }
            }
        }
    }
}

/* gx (nullable) must hold B*C*H*W floats, it is zeroed here; gflow nullable. */
void usf_oracle_warp_bwd(const float *x, const float *flow, const float *gout, float *gx,
                         float *gflow, int B, int C, int H, int W, int pad_mode) {
    const size_t plane = (size_t)H * W;
    double *gacc = NULL;
    if (gx) gacc = (double *)calloc((size_t)B * C * plane, sizeof(double));
    /* parallel over b only: the scatter into gacc stays race-free */
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int y = 0; y < H; ++y) {
            for (int xo = 0; xo < W; ++xo) {
                const float u = flow[IDX4(b, 0, y, xo, 2, H, W)];
                const float v = flow[IDX4(b, 1, y, xo, 2, H, W)];
                float mx, my;
                const float ix = src_coord(xo, u, W, pad_mode, &mx);
                const float iy = src_coord(y, v, H, pad_mode, &my);
                const float fx = floorf(ix), fy = floorf(iy);
                const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
                const float ex = (fx + 1.0f) - ix, wx = ix - fx;
                const float ey = (fy + 1.0f) - iy, wy = iy - fy;
                const float nw = ex * ey, ne = wx * ey, sw = ex * wy, se = wx * wy;
                const int i_nw = inb(y0, x0, H, W), i_ne = inb(y0, x1, H, W);
                const int i_sw = inb(y1, x0, H, W), i_se = inb(y1, x1, H, W);
                double gix = 0.0, giy = 0.0;
                for (int c = 0; c < C; ++c) {
                    const float go = gout[IDX4(b, c, y, xo, C, H, W)];
                    const float *p = x + IDX4(b, c, 0, 0, C, H, W);
                    if (gacc) {
                        double *q = gacc + ((size_t)b * C + c) * plane;
                        if (i_nw) q[(size_t)y0 * W + x0] += (double)(nw * go);
                        if (i_ne) q[(size_t)y0 * W + x1] += (double)(ne * go);
                        if (i_sw) q[(size_t)y1 * W + x0] += (double)(sw * go);
                        if (i_se) q[(size_t)y1 * W + x1] += (double)(se * go);
                    }
                    if (gflow) {
                        if (i_nw) {
                            const float val = p[(size_t)y0 * W + x0];
                            gix -= (double)(val * ey * go);
                            giy -= (double)(val * ex * go);
                        }
                        if (i_ne) {
                            const float val = p[(size_t)y0 * W + x1];
                            gix += (double)(val * ey * go);
                            giy -= (double)(val * wx * go);
                        }
                        if (i_sw) {
                            const float val = p[(size_t)y1 * W + x0];
                            gix -= (double)(val * wy * go);
                            giy += (double)(val * ex * go);
                        }
                        if (i_se) {
                            const float val = p[(size_t)y1 * W + x1];
                            gix += (double)(val * wy * go);
                            giy += (double)(val * wx * go);
                        }
                    }
                }
                if (gflow) {
                    /* grid_sampler backward: ggrid = mult * gi; norm_grid backward:
                     * d(2.0*v/(size-1) - 1.0) -> (g / (size-1)) * 2.0 */
                    const float gxn = mx * (float)gix, gyn = my * (float)giy;
                    gflow[IDX4(b, 0, y, xo, 2, H, W)] = (gxn / (float)(W - 1)) * 2.0f;
                    gflow[IDX4(b, 1, y, xo, 2, H, W)] = (gyn / (float)(H - 1)) * 2.0f;
                }
            }
        }
    }
    if (gacc) {
        const size_t n = (size_t)B * C * plane;
        for (size_t i = 0; i < n; ++i) gx[i] = (float)gacc[i];
        free(gacc);
    }
}

/* ------------------------------------------------------------------------------------------
 * Occlusion masks.
 *
 *   usf_oracle_occ_bidirection  utils/warp_utils.py:109-117 get_occu_mask_bidirection:
 *       w = flow_warp(flow21, flow12, pad="zeros"); diff = flow12 + w;
 *       mag = sum_c flow12^2 + sum_c w^2; occ = sum_c diff^2 > scale*mag + bias.
 *       Channel sums have two terms: one rounded fp32 add each.  margin (nullable)
 *       receives lhs - threshold in double, so a test can tell a genuine mismatch
 *       from a pixel sitting on the threshold.
 *   usf_oracle_occ_backward     utils/warp_utils.py:120-126 get_occu_mask_backward with
 *       get_corresponding_map (:26-94) on mesh_grid + flow21: every source pixel adds
 *       (1-|x-xn|)*(1-|y-yn|) to each of its four integer neighbours (xn, yn) that lie
 *       inside the image (the reference zeroes the taps whose clamped index differs from
 *       the unclamped one, :52-65,89); occ = clamp(range, 0, 1) < th.  The weights are
 *       separately rounded fp32 operations as in the reference; the range map accumulates
 *       in double (the reference's scatter_add_ sums in fp32 in source order).
 *       Non-finite coordinates contribute nothing (undefined in the reference: .long()).
 */
void usf_oracle_occ_bidirection(const float *flow12, const float *flow21, float *occ, double *margin,
                                int B, int H, int W, float scale, float bias) {
    const size_t plane = (size_t)H * W;
    float *warped = (float *)malloc(sizeof(float) * 2 * plane * (size_t)B);
    usf_oracle_warp_fwd(flow21, flow12, warped, B, 2, H, W, 1 /* zeros */);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        for (size_t p = 0; p < plane; ++p) {
            const float u = flow12[((size_t)b * 2) * plane + p], v = flow12[((size_t)b * 2 + 1) * plane + p];
            const float wu = warped[((size_t)b * 2) * plane + p], wv = warped[((size_t)b * 2 + 1) * plane + p];
            volatile float d0 = u + wu, d1 = v + wv;
            volatile float uu = u * u, vv = v * v, m12 = uu + vv;
            volatile float w0 = wu * wu, w1 = wv * wv, mw = w0 + w1;
            volatile float mag = m12 + mw;
            volatile float sm = scale * mag;
            volatile float thresh = sm + bias;
            volatile float q0 = d0 * d0, q1 = d1 * d1, lhs = q0 + q1;
            occ[(size_t)b * plane + p] = lhs > thresh ? 1.0f : 0.0f;
            if (margin) margin[(size_t)b * plane + p] = (double)lhs - (double)thresh;
        }
    }
    free(warped);
}

/* One neighbour of get_corresponding_map: weight (1-|x-xn|)*(1-|y-yn|), zero when outside. */
static inline void splat_tap(double *acc, float x, float y, float xn, float yn, int H, int W) {
    if (!(xn >= 0.0f && xn <= (float)(W - 1) && yn >= 0.0f && yn <= (float)(H - 1))) return;
    volatile float ax = x - xn, ay = y - yn;
    volatile float wx = 1.0f - fabsf(ax), wy = 1.0f - fabsf(ay);
    volatile float w = wx * wy;
    acc[(size_t)yn * W + (size_t)xn] += (double)w;
}

void usf_oracle_occ_backward(const float *flow21, float *occ, float *range_map, int B, int H, int W,
                             float th) {
    const size_t plane = (size_t)H * W;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        double *acc = (double *)calloc(plane, sizeof(double));
        for (int yo = 0; yo < H; ++yo) {
            for (int xo = 0; xo < W; ++xo) {
                volatile float x = (float)xo + flow21[IDX4(b, 0, yo, xo, 2, H, W)];
                volatile float y = (float)yo + flow21[IDX4(b, 1, yo, xo, 2, H, W)];
                if (!isfinite(x) || !isfinite(y)) continue;
                const float x1 = floorf(x), y1 = floorf(y);
                volatile float x0 = x1 + 1.0f, y0 = y1 + 1.0f;
                splat_tap(acc, x, y, x0, y0, H, W); /* ceil, ceil */
                splat_tap(acc, x, y, x0, y1, H, W); /* ceil x, floor y */
                splat_tap(acc, x, y, x1, y0, H, W); /* floor x, ceil y */
                splat_tap(acc, x, y, x1, y1, H, W); /* floor, floor */
            }
        }
        for (size_t p = 0; p < plane; ++p) {
            const float r = (float)acc[p];
            if (range_map) range_map[(size_t)b * plane + p] = r;
            const float c = r < 0.0f ? 0.0f : (r > 1.0f ? 1.0f : r);
            occ[(size_t)b * plane + p] = c < th ? 1.0f : 0.0f;
        }
        free(acc);
    }
}
