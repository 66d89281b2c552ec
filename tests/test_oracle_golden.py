"""Pins oracle/ (C restatement + torch port) against outputs of the reference itself.

Fixtures: tests/golden/*.npz, produced by oracle/make_golden.py from the unmodified
reference.  Tolerance: the oracle accumulates in double, the reference in fp32, so
they differ by fp32 summation noise only (<= 2e-6 relative here; parity budget 1e-5).
"""
import numpy as np
import pytest
import torch

from conftest import max_norm_err, rel_l2
from oracle import c_oracle, torch_port

TOL = 2e-6


def _check(got, want, what):
    assert got.shape == want.shape, what
    assert rel_l2(got, want) <= TOL, (what, rel_l2(got, want))
    assert max_norm_err(got, want) <= TOL, (what, max_norm_err(got, want))


def test_c_oracle_corr_matches_reference(corr_golden):
    assert len(corr_golden) >= 7
    for name, t in corr_golden.items():
        _check(c_oracle.corr_fwd(t["x1"], t["x2"]), t["out"], f"{name}.out")
        gx1, gx2 = c_oracle.corr_bwd(t["x1"], t["x2"], t["gout"])
        _check(gx1, t["gx1"], f"{name}.gx1")
        _check(gx2, t["gx2"], f"{name}.gx2")


def test_c_oracle_corr_bwd_nullable(corr_golden):
    t = corr_golden["tiny"]
    gx1, none = c_oracle.corr_bwd(t["x1"], t["x2"], t["gout"], need_gx2=False)
    assert none is None
    _check(gx1, t["gx1"], "gx1 only")


def test_c_oracle_warp_matches_reference(warp_golden):
    assert len(warp_golden) >= 10
    for name, t in warp_golden.items():
        pad = str(t["pad"])
        _check(c_oracle.warp_fwd(t["x"], t["flow"], pad), t["out"], f"{name}.out")
        gx, gflow = c_oracle.warp_bwd(t["x"], t["flow"], t["gout"], pad)
        _check(gx, t["gx"], f"{name}.gx")
        _check(gflow, t["gflow"], f"{name}.gflow")


def test_torch_port_matches_reference(corr_golden, warp_golden):
    torch.set_num_threads(1)
    for name, t in corr_golden.items():
        out, gx1, gx2 = torch_port.corr_fwd_bwd(*(torch.from_numpy(t[k]) for k in ("x1", "x2", "gout")))
        # same ATen ops, same order, same thread count -> bit-identical
        assert np.array_equal(out.numpy(), t["out"]), name
        assert np.array_equal(gx1.numpy(), t["gx1"]), name
        assert np.array_equal(gx2.numpy(), t["gx2"]), name
    for name, t in warp_golden.items():
        out, gx, gflow = torch_port.warp_fwd_bwd(
            *(torch.from_numpy(t[k]) for k in ("x", "flow", "gout")), pad=str(t["pad"]))
        assert np.array_equal(out.numpy(), t["out"]), name
        _check(gx.numpy(), t["gx"], f"{name}.gx")  # scatter order may differ
        assert np.array_equal(gflow.numpy(), t["gflow"]), name


def test_oracle_zero_flow_is_identity(warp_golden):
    t = warp_golden["border_zero_flow"]
    out = c_oracle.warp_fwd(t["x"], t["flow"], "border")
    # x + 0 survives the normalise/unnormalise round trip to ~1 ulp of W
    assert max_norm_err(out, t["x"]) < 1e-5


def test_oracle_threads_report():
    assert c_oracle.num_threads() >= 1


# ---- occlusion masks (utils/warp_utils.py:109-126) -------------------------------------------------
# The masks are comparisons: a pixel whose two sides differ by summation noise may flip between two
# correct fp32 evaluations.  A mismatch is accepted only where the oracle's own margin is at that level.
OCC_MARGIN = 1e-5


def occ_mismatch(got, want, margin, rel=None):
    """Number of pixels where the masks differ although the oracle's margin is NOT at rounding level."""
    scale = 1.0 if rel is None else np.maximum(np.abs(rel), 1.0)
    return int(((got != want) & (np.abs(margin) > OCC_MARGIN * scale)).sum())


def test_c_oracle_occ_masks_match_reference(occ_golden):
    assert len(occ_golden) >= 7
    flips = 0
    for name, t in occ_golden.items():
        occ, margin = c_oracle.occ_bidirection(t["flow12"], t["flow21"])
        assert occ.shape == t["occ_bidirection"].shape, name
        assert set(np.unique(occ)) <= {0.0, 1.0}, name
        assert occ_mismatch(occ, t["occ_bidirection"], margin) == 0, name
        flips += int((occ != t["occ_bidirection"]).sum())
        occ, rng = c_oracle.occ_backward(t["flow21"])
        _check(rng, t["range_map"], f"{name}.range_map")
        assert occ_mismatch(occ, t["occ_backward"], t["range_map"] - np.float32(0.2)) == 0, name
        flips += int((occ != t["occ_backward"]).sum())
    assert flips <= 2, flips  # on these fixtures the oracle is in fact bit-identical or nearly so


def test_torch_port_occ_masks_match_reference(occ_golden):
    torch.set_num_threads(1)
    for name, t in occ_golden.items():
        f12, f21 = torch.from_numpy(t["flow12"]), torch.from_numpy(t["flow21"])
        assert np.array_equal(torch_port.occ_mask_bidirection(f12, f21).numpy(), t["occ_bidirection"]), name
        assert np.array_equal(torch_port.occ_mask_backward(f21).numpy(), t["occ_backward"]), name
        base = torch_port.pixel_grid(*f21.shape[:1], *f21.shape[2:]).type_as(f21)
        assert np.array_equal(torch_port.range_map(base + f21).numpy(), t["range_map"]), name


def test_oracle_occ_backward_properties():
    # zero flow: every pixel is hit by exactly its own weight 1 -> nothing occluded
    occ, rng = c_oracle.occ_backward(np.zeros((2, 2, 5, 7), np.float32))
    assert np.array_equal(rng, np.ones_like(rng)) and not occ.any()
    # everything leaves the image -> everything occluded, empty range map
    occ, rng = c_oracle.occ_backward(np.full((1, 2, 5, 7), 100.0, np.float32))
    assert occ.all() and not rng.any()
    # mass conservation: with all taps inside, the range map sums to the number of pixels
    rs = np.random.RandomState(0)
    flow = rs.uniform(-0.4, 0.4, (1, 2, 12, 12)).astype(np.float32)
    flow[:, :, 0, :] = flow[:, :, -1, :] = 0
    flow[:, :, :, 0] = flow[:, :, :, -1] = 0
    _, rng = c_oracle.occ_backward(flow)
    assert abs(float(rng.astype(np.float64).sum()) - 144.0) < 1e-4
    # non-finite coordinates contribute nothing
    flow[0, 0, 3, 3] = np.nan
    flow[0, 1, 4, 4] = np.inf
    _, rng2 = c_oracle.occ_backward(flow)
    assert np.isfinite(rng2).all() and abs(float(rng2.astype(np.float64).sum()) - 142.0) < 1e-4


# ---- size-independent properties of the oracle itself (ragged shapes the fixtures do not hold) ---------------------
@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (2, 3, 2, 17), (1, 33, 11, 5), (3, 4, 9, 9), (1, 196, 3, 10)])
def test_oracle_corr_adjoint_identities(shape):
    """The cost volume is bilinear in (x1, x2):  <out, g> = <x1, gx1> = <x2, gx2>."""
    rs = np.random.RandomState(sum(shape))
    B, C, H, W = shape
    x1, x2 = rs.randn(*shape).astype(np.float32), rs.randn(*shape).astype(np.float32)
    g = rs.randn(B, 81, H, W).astype(np.float32)
    out = c_oracle.corr_fwd(x1, x2).astype(np.float64)
    gx1, gx2 = c_oracle.corr_bwd(x1, x2, g)
    lhs = float((out * g).sum())
    scale = float(np.abs(out * g).sum()) + 1e-30
    assert abs(lhs - float((x1.astype(np.float64) * gx1).sum())) <= 1e-5 * scale
    assert abs(lhs - float((x2.astype(np.float64) * gx2).sum())) <= 1e-5 * scale


@pytest.mark.parametrize("pad", ["border", "zeros"])
@pytest.mark.parametrize("shape", [(1, 1, 2, 2), (2, 3, 5, 17), (1, 7, 13, 4)])
def test_oracle_warp_is_linear_in_x_with_adjoint_gx(shape, pad):
    """flow_warp is linear in x for a fixed flow:  <warp(x), g> = <x, gx>, and gx does not depend on x."""
    rs = np.random.RandomState(sum(shape) + len(pad))
    B, C, H, W = shape
    x = rs.randn(*shape).astype(np.float32)
    flow = (2.5 * rs.randn(B, 2, H, W)).astype(np.float32)
    g = rs.randn(*shape).astype(np.float32)
    out = c_oracle.warp_fwd(x, flow, pad).astype(np.float64)
    gx, _ = c_oracle.warp_bwd(x, flow, g, pad)
    lhs = float((out * g).sum())
    scale = float(np.abs(out * g).sum()) + 1e-30
    assert abs(lhs - float((x.astype(np.float64) * gx).sum())) <= 1e-5 * scale
    gx_other, _ = c_oracle.warp_bwd(rs.randn(*shape).astype(np.float32), flow, g, pad)
    assert np.array_equal(gx, gx_other)
