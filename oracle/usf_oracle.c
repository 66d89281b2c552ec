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
