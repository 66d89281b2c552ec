"""numpy front end of oracle/usf_oracle.c (TEST INFRASTRUCTURE ONLY).

Each function cites the reference code it restates; see the C file's header for
the full map.  Arrays are float32, NCHW, C-contiguous.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libusf_oracle.so")
_lib = None

PAD_MODES = {"border": 0, "zeros": 1}


def build(force=False):
    """Compile usf_oracle.c with the committed Makefile (gcc, -ffp-contract=off)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "usf_oracle.c")
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        i = ctypes.c_int
        L.usf_oracle_corr_fwd.argtypes = [fp, fp, fp, i, i, i, i, i]
        L.usf_oracle_corr_bwd.argtypes = [fp, fp, fp, fp, fp, i, i, i, i, i]
        L.usf_oracle_warp_fwd.argtypes = [fp, fp, fp, i, i, i, i, i]
        L.usf_oracle_warp_bwd.argtypes = [fp, fp, fp, fp, fp, i, i, i, i, i]
        dp = ctypes.POINTER(ctypes.c_double)
        L.usf_oracle_occ_bidirection.argtypes = [fp, fp, fp, dp, i, i, i, ctypes.c_float, ctypes.c_float]
        L.usf_oracle_occ_backward.argtypes = [fp, fp, fp, i, i, i, ctypes.c_float]
        L.usf_oracle_occ_bidirection.restype = None
        L.usf_oracle_occ_backward.restype = None
        L.usf_oracle_num_threads.restype = i
        L.usf_oracle_set_num_threads.argtypes = [i]
        for f in (L.usf_oracle_corr_fwd, L.usf_oracle_corr_bwd, L.usf_oracle_warp_fwd,
                  L.usf_oracle_warp_bwd, L.usf_oracle_set_num_threads):
            f.restype = None
        _lib = L
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def num_threads():
    return int(lib().usf_oracle_num_threads())


def set_num_threads(n):
    lib().usf_oracle_set_num_threads(int(n))


def corr_fwd(x1, x2, max_displacement=4):
    """models/correlation_native.py:13-23 -> [B, (2md+1)^2, H, W]."""
    x1, p1 = _f32(x1)
    x2, p2 = _f32(x2)
    assert x1.shape == x2.shape and x1.ndim == 4
    B, C, H, W = x1.shape
    D = 2 * max_displacement + 1
    out = np.empty((B, D * D, H, W), dtype=np.float32)
    lib().usf_oracle_corr_fwd(p1, p2, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), B, C, H, W,
                              max_displacement)
    return out


def corr_bwd(x1, x2, gout, max_displacement=4, need_gx1=True, need_gx2=True):
    """autograd of correlation_native.Correlation.forward -> (gx1, gx2)."""
    x1, p1 = _f32(x1)
    x2, p2 = _f32(x2)
    gout, pg = _f32(gout)
    B, C, H, W = x1.shape
    D = 2 * max_displacement + 1
    assert gout.shape == (B, D * D, H, W)
    null = ctypes.POINTER(ctypes.c_float)()
    gx1 = np.empty_like(x1) if need_gx1 else None
    gx2 = np.empty_like(x2) if need_gx2 else None
    lib().usf_oracle_corr_bwd(
        p1, p2, pg,
        gx1.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if need_gx1 else null,
        gx2.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if need_gx2 else null,
        B, C, H, W, max_displacement)
    return gx1, gx2


def warp_fwd(x, flow, pad="border"):
    """utils/warp_utils.py:97-106 flow_warp(x, flow12, pad, mode="bilinear")."""
    x, px = _f32(x)
    flow, pf = _f32(flow)
    B, C, H, W = x.shape
    assert flow.shape == (B, 2, H, W)
    out = np.empty_like(x)
    lib().usf_oracle_warp_fwd(px, pf, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), B, C, H, W,
                              PAD_MODES[pad])
    return out


def warp_bwd(x, flow, gout, pad="border", need_gx=True, need_gflow=True):
    """autograd of flow_warp wrt x and flow12 -> (gx, gflow)."""
    x, px = _f32(x)
    flow, pf = _f32(flow)
    gout, pg = _f32(gout)
    B, C, H, W = x.shape
    null = ctypes.POINTER(ctypes.c_float)()
    gx = np.empty_like(x) if need_gx else None
    gflow = np.empty_like(flow) if need_gflow else None
    lib().usf_oracle_warp_bwd(
        px, pf, pg,
        gx.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if need_gx else null,
        gflow.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if need_gflow else null,
        B, C, H, W, PAD_MODES[pad])
    return gx, gflow


def occ_bidirection(flow12, flow21, scale=0.01, bias=0.5):
    """utils/warp_utils.py:109-117 get_occu_mask_bidirection -> (occ [B,1,H,W], margin float64 [B,1,H,W]);
    margin = lhs - threshold: a pixel whose |margin| is at rounding level may legitimately flip."""
    flow12, p12 = _f32(flow12)
    flow21, p21 = _f32(flow21)
    B, two, H, W = flow12.shape
    assert two == 2 and flow21.shape == flow12.shape
    occ = np.empty((B, 1, H, W), dtype=np.float32)
    margin = np.empty((B, 1, H, W), dtype=np.float64)
    lib().usf_oracle_occ_bidirection(p12, p21, occ.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                     margin.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), B, H, W,
                                     scale, bias)
    return occ, margin


def occ_backward(flow21, th=0.2):
    """utils/warp_utils.py:120-126 get_occu_mask_backward -> (occ [B,1,H,W], range map [B,1,H,W] =
    get_corresponding_map's unclamped return value, :26-94)."""
    flow21, p21 = _f32(flow21)
    B, two, H, W = flow21.shape
    assert two == 2
    occ = np.empty((B, 1, H, W), dtype=np.float32)
    rng = np.empty((B, 1, H, W), dtype=np.float32)
    lib().usf_oracle_occ_backward(p21, occ.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                  rng.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), B, H, W, th)
    return occ, rng
