"""The C-ABI library loads and exports every symbol include/unsamflow_b200.h declares.
No GPU needed: only argument validation paths are called (they return before any CUDA call)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "unsamflow_b200.h")


@pytest.fixture(scope="module")
def lib():
    from unsamflow_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(usf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from unsamflow_b200 import _lib
    names = declared_symbols()
    assert len(names) == 18
    assert sorted(_lib.SYMBOLS) == names
    for n in names:
        assert getattr(lib, n) is not None


def test_abi_version_and_status_strings(lib):
    assert lib.usf_abi_version() == 3
    for st in (0, -1, -2, -3, -4, -5):
        assert lib.usf_status_string(st)
    assert b"unknown" in lib.usf_status_string(-99)


def test_argument_validation_without_a_gpu(lib):
    buf = (ctypes.c_float * 16)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    BAD_SHAPE, BAD_ARG, WORKSPACE, UNSUPPORTED = -1, -2, -3, -5
    # null pointers / wrong displacement / unknown flag / unknown dtype
    assert lib.usf_corr_fwd(None, p, p, 1, 1, 2, 2, 4, 0, 0, 0.0, None) == BAD_ARG
    assert lib.usf_corr_fwd(p, p, p, 1, 1, 2, 2, 3, 0, 0, 0.0, None) == BAD_ARG
    assert lib.usf_corr_fwd(p, p, p, 1, 1, 2, 2, 4, 0, 64, 0.0, None) == BAD_ARG
    assert lib.usf_corr_fwd(p, p, p, 1, 1, 2, 2, 4, 7, 0, 0.0, None) == BAD_ARG
    # shapes: non-positive, and more than 2^31-1 elements
    assert lib.usf_corr_fwd(p, p, p, 0, 1, 2, 2, 4, 0, 0, 0.0, None) == BAD_SHAPE
    assert lib.usf_corr_fwd(p, p, p, 64, 512, 1024, 1024, 4, 0, 0, 0.0, None) == BAD_SHAPE
    assert lib.usf_corr_bwd(p, p, None, p, p, 1, 1, 2, 2, 4, 0, 0, None) == BAD_ARG
    assert lib.usf_corr_bwd(p, p, p, None, None, 1, 1, 2, 2, 4, 0, 0, None) == 0  # nothing requested
    # warp: pad mode, H/W >= 2 (the reference divides by size-1), workspace
    assert lib.usf_warp_fwd(p, p, p, 1, 1, 2, 2, 5, 0, 0, None) == BAD_ARG
    assert lib.usf_warp_fwd(p, p, p, 1, 1, 1, 2, 0, 0, 0, None) == BAD_SHAPE
    assert lib.usf_warp_bwd(p, p, p, None, None, 1, 1, 2, 2, 0, 0, 0, None, 0, None) == 0
    # images above 1024 pixels take the three-kernel path, which needs the caller's workspace
    assert lib.usf_warp_bwd(p, p, p, p, p, 1, 1, 40, 40, 0, 0, 0, None, 0, None) == WORKSPACE
    need = lib.usf_warp_bwd_workspace_bytes(1, 1, 40, 40, 0)
    assert need > 0
    assert lib.usf_warp_bwd(p, p, p, p, p, 1, 1, 40, 40, 0, 0, 0, p, need - 1, None) == WORKSPACE
    assert lib.usf_warp_bwd_workspace_bytes(1, 1, 40, 40, 4) == 0  # atomic mode needs none
    assert lib.usf_warp_bwd_workspace_bytes(8, 32, 64, 208, 0) == 8 * 64 * 208 * (4 + 64 + 48) + 16
    assert lib.usf_warp_bwd_workspace_bytes(1, 1, 33, 33, 0) == 4368 + 16 + 1089 * (64 + 48)  # cnt padded to 16 B
    # ... smaller ones run as one kernel with shared-memory tap lists: no workspace
    assert lib.usf_warp_bwd_workspace_bytes(8, 96, 16, 52, 0) == 0
    assert lib.usf_warp_bwd_workspace_bytes(1, 1, 32, 32, 0) == 0
    # the counter region a USF_FLAG_WORKSPACE_CLEAN caller zeroes once: cnt (padded to 16 B) + 16
    assert lib.usf_warp_bwd_workspace_clean_bytes(8, 32, 64, 208, 0) == 8 * 64 * 208 * 4 + 16
    assert lib.usf_warp_bwd_workspace_clean_bytes(1, 1, 33, 33, 0) == 4368 + 16
    assert lib.usf_warp_bwd_workspace_clean_bytes(8, 96, 16, 52, 0) == 0
    assert lib.usf_warp_bwd_workspace_clean_bytes(1, 1, 40, 40, 4) == 0
    # occlusion masks: null pointers, unknown flag, H/W >= 2 for the warp inside, workspace size / alignment
    assert lib.usf_occ_bidirection(p, None, p, 1, 2, 2, 0.01, 0.5, 0, None) == BAD_ARG
    assert lib.usf_occ_bidirection(p, p, p, 1, 2, 2, 0.01, 0.5, 4, None) == BAD_ARG
    assert lib.usf_occ_bidirection(p, p, p, 1, 1, 2, 0.01, 0.5, 0, None) == BAD_SHAPE
    assert lib.usf_occ_backward_workspace_bytes(8, 256, 832) == 8 * 256 * 832 * 8
    assert lib.usf_occ_backward_workspace_bytes(0, 4, 4) == 0
    assert lib.usf_occ_backward(None, p, None, 1, 2, 2, 0.2, p, 32, None) == BAD_ARG
    assert lib.usf_occ_backward(p, p, None, 1, 0, 2, 0.2, p, 32, None) == BAD_SHAPE
    assert lib.usf_occ_backward(p, p, None, 1, 2, 2, 0.2, None, 0, None) == WORKSPACE
    assert lib.usf_occ_backward(p, p, None, 1, 2, 2, 0.2, p, 31, None) == WORKSPACE
    assert lib.usf_occ_backward(p, p, None, 1, 2, 2, 0.2, ctypes.c_void_p(p.value + 4), 32, None) == WORKSPACE


def test_launch_counter_is_per_thread_and_resettable(lib):
    lib.usf_reset_launch_count()
    assert lib.usf_launch_count() == 0
    assert lib.usf_last_cuda_error() == 0
