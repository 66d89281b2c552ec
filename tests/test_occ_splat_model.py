"""Bit-exact Python model of the range-map splat's device algorithm (unsamflow_b200/csrc/occ.cu) against the oracle.

The kernel cannot run here (no GPU), but its correctness argument can: every variant -- four atomics per pixel,
the lane-merged form (lane i+1 folds lane i's ceil-x taps into its own floor-x taps when the target coordinates
coincide) and the shared-memory window form (per 16x64 source tile, flushed once) -- only regroups INTEGER adds of
the same 2^-40 fixed-point terms.  This model restates the index logic of those variants (clamped floor, validity
from the unclamped float, window origin from the tile's centre flow) and checks on ragged / wild / non-finite
flows that (a) no variant ever forms an address outside the image, (b) all variants produce identical integers,
(c) the integers, converted with one rounding, equal the oracle's range map to the last bit on these cases.
The GPU tests (tests/test_gpu_occlusion.py) check the same properties on the device.
"""
import numpy as np
import pytest

from oracle import c_oracle

f32 = np.float32
SCALE = 2 ** 40


def axis(i, fl, size):
    """splat_axis(): floor neighbour (clamped to [-1, size-1]), validity of floor / ceil, their weights."""
    with np.errstate(all="ignore"):
        x = f32(f32(i) + f32(fl))
        x1 = np.floor(x)
        x0 = f32(x1 + f32(1))
        mx = f32(size - 1)
        lo_ok = bool(x1 >= 0 and x1 <= mx)
        hi_ok = bool(x0 >= 0 and x0 <= mx)
        t = x1 if not np.isnan(x1) else f32(-1)      # fmaxf(NaN, -1) = -1
        lo = int(min(max(t, f32(-1)), mx))
        w_lo = f32(f32(1) - abs(f32(x - x1)))
        w_hi = f32(f32(1) - abs(f32(x - x0)))
    return lo, lo_ok, hi_ok, w_lo, w_hi


def fix(ok, w):
    return int(round(float(w) * SCALE)) if (ok and w > 0) else 0   # w * 2^40 is exact in fp32: no tie cases


def taps_of(flow, b, y, xo):
    _, _, H, W = flow.shape
    ax, ay = axis(xo, flow[b, 0, y, xo], W), axis(y, flow[b, 1, y, xo], H)
    return ax[0], ay[0], [fix(ax[1] and ay[1], f32(ax[3] * ay[3])), fix(ax[2] and ay[1], f32(ax[4] * ay[3])),
                          fix(ax[1] and ay[2], f32(ax[3] * ay[4])), fix(ax[2] and ay[2], f32(ax[4] * ay[4]))]


def splat_lanes(flow, merge):
    """occ_splat_kernel<MERGE, 1>: one pixel per lane, warps of 32 consecutive pixels."""
    B, _, H, W = flow.shape
    n = B * H * W
    acc = np.zeros((B, H * W), dtype=object)
    acc[:] = 0
    for w0 in range(0, (n + 31) // 32 * 32, 32):
        lanes = []
        for lane in range(32):
            p = w0 + lane
            if p >= n:
                lanes.append([False, -1, -1, -1, [0, 0, 0, 0]])
                continue
            b, rem = divmod(p, H * W)
            xlo, ylo, t = taps_of(flow, b, rem // W, rem % W)
            lanes.append([True, b, xlo, ylo, t])
        if merge:
            take = [False] + [lanes[i][0] and lanes[i - 1][1] == lanes[i][1] and lanes[i - 1][2] + 1 == lanes[i][2]
                              and lanes[i - 1][3] == lanes[i][3] for i in range(1, 32)]
            before = [list(l[4]) for l in lanes]
            for i in range(32):
                if take[i]:
                    lanes[i][4][0] += before[i - 1][1]
                    lanes[i][4][2] += before[i - 1][3]
                if i < 31 and take[i + 1]:
                    lanes[i][4][1] = lanes[i][4][3] = 0
        for live, b, xlo, ylo, t in lanes:
            if live:
                for k, off in enumerate((0, 1, W, W + 1)):
                    if t[k]:
                        assert 0 <= ylo * W + xlo + off < H * W
                        acc[b, ylo * W + xlo + off] += t[k]
    return acc


def splat_tiles(flow, th=16, tw=64, r=8):
    """occ_splat_tile_kernel: per-tile shared-memory window + flush; out-of-window taps go straight to global."""
    B, _, H, W = flow.shape
    wh, ww = th + 2 * r + 1, tw + 2 * r + 1
    acc = np.zeros((B, H * W), dtype=object)
    acc[:] = 0
    tiles_x, tiles_y = (W + tw - 1) // tw, (H + th - 1) // th
    outside = 0
    for blk in range(B * tiles_x * tiles_y):
        t = blk
        tx0 = (t % tiles_x) * tw
        t //= tiles_x
        ty0 = (t % tiles_y) * th
        b = t // tiles_y
        cu, cv = flow[b, 0, min(ty0 + th // 2, H - 1), min(tx0 + tw // 2, W - 1)], \
            flow[b, 1, min(ty0 + th // 2, H - 1), min(tx0 + tw // 2, W - 1)]
        ou = int(np.rint(cu)) if abs(cu) < 1048576 else 0     # NaN compares false -> 0
        ov = int(np.rint(cv)) if abs(cv) < 1048576 else 0
        wx0, wy0 = tx0 + ou - r, ty0 + ov - r
        win = [0] * (wh * ww)
        for y in range(ty0, min(ty0 + th, H)):
            for xo in range(tx0, min(tx0 + tw, W)):
                xlo, ylo, taps = taps_of(flow, b, y, xo)
                for j, v in enumerate(taps):
                    if v:
                        X, Y = xlo + (j & 1), ylo + (j >> 1)
                        assert 0 <= X < W and 0 <= Y < H
                        rx, ry = X - wx0, Y - wy0
                        if 0 <= rx < ww and 0 <= ry < wh:
                            win[ry * ww + rx] += v
                        else:
                            acc[b, Y * W + X] += v
                            outside += 1
        for i, s in enumerate(win):
            if s:
                ry, rx = divmod(i, ww)
                assert 0 <= wx0 + rx < W and 0 <= wy0 + ry < H
                acc[b, (wy0 + ry) * W + wx0 + rx] += s
    return acc, outside


def flows():
    rs = np.random.RandomState(0)
    ramp = np.tile(np.linspace(-2, 2, 45)[None, None, None, :], (2, 2, 7, 1)) + 0.3
    ys, xs = np.meshgrid(np.arange(7.0), np.arange(9.0), indexing="ij")
    bad = rs.randn(1, 2, 6, 7)
    bad[0, 0, 2, 2], bad[0, 1, 3, 3], bad[0, 0, 4, 4], bad[0, 1, 1, 1], bad[0, 0, 0, 0] = np.nan, np.inf, -np.inf, 3e38, -1e30
    patch = np.tile(np.linspace(-3, 3, 70)[None, None, None, :], (1, 2, 20, 1))
    patch[:, :, 5:12, 20:50] += 25.0 * rs.randn(1, 2, 7, 30)   # leaves the tile window
    return {"gauss": 3 * rs.randn(2, 2, 9, 11), "ramp": ramp, "wild": rs.randn(1, 2, 8, 26) * 8,
            "integer": rs.randint(-3, 4, (1, 2, 6, 13)).astype(float), "collapse": np.stack([2.25 - xs, 3.5 - ys])[None],
            "non_finite": bad, "patch": patch}


@pytest.mark.parametrize("name", sorted(flows()))
def test_splat_variants_regroup_the_same_integer_sum(name):
    flow = np.ascontiguousarray(flows()[name], dtype=f32)
    plain = splat_lanes(flow, merge=False)
    merged = splat_lanes(flow, merge=True)
    tiled, outside = splat_tiles(flow)
    assert (plain == merged).all() and (plain == tiled).all()
    if name == "patch":
        assert outside > 0          # the global path of the window variant is exercised
    B, _, H, W = flow.shape
    rng = np.array([[f32(float(v) / SCALE) for v in row] for row in plain], dtype=f32).reshape(B, 1, H, W)
    _, want = c_oracle.occ_backward(flow)
    assert np.array_equal(rng, want), float(np.abs(rng - want).max())
