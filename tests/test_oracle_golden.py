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
