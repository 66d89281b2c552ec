"""ctypes binding of libunsamflow_b200.so (the C ABI in include/unsamflow_b200.h).

There is NO fallback: if the CUDA library is missing or fails to load, importing the
ops raises.  Nothing here touches oracle/.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# USF_LIB_PATH overrides the library location (A/B runs against another build of the same ABI)
LIB_PATH = os.environ.get("USF_LIB_PATH") or os.path.join(_PKG, "lib", "libunsamflow_b200.so")

ABI_VERSION = 3
F32, BF16 = 0, 1
PAD_BORDER, PAD_ZEROS = 0, 1
FLAG_LEAKY_RELU = 1
FLAG_COORD_RECIPROCAL = 2
FLAG_WARP_BWD_ATOMIC = 4
FLAG_WORKSPACE_CLEAN = 8

# every symbol include/unsamflow_b200.h declares (tests/test_abi.py checks header <-> library <-> this list)
SYMBOLS = (
    "usf_abi_version", "usf_status_string", "usf_last_cuda_error", "usf_launch_count",
    "usf_reset_launch_count", "usf_corr_fwd", "usf_corr_fwd_into", "usf_corr_bwd", "usf_warp_fwd",
    "usf_warp_bwd_workspace_bytes", "usf_warp_bwd_workspace_clean_bytes", "usf_warp_bwd", "usf_occ_bidirection",
    "usf_occ_backward_workspace_bytes", "usf_occ_backward", "usf_segpool_max_segments", "usf_segpool_fwd",
    "usf_segpool_bwd",
)

_lib = None


class UsfError(RuntimeError):
    """Raised when a C-ABI call returns non-zero (the reference raises RuntimeError via AT_ERROR,
    models/correlation_package/correlation_cuda.cc:80-82)."""


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m unsamflow_b200.build` "
            "(or __graft_entry__.build()). unsamflow_b200 has no CPU or PyTorch fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i, u, f, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_float, ctypes.c_size_t
    L.usf_abi_version.restype = i
    L.usf_status_string.restype = ctypes.c_char_p
    L.usf_status_string.argtypes = [i]
    L.usf_last_cuda_error.restype = i
    L.usf_launch_count.restype = ctypes.c_ulonglong
    L.usf_reset_launch_count.restype = None
    L.usf_corr_fwd.restype = i
    L.usf_corr_fwd.argtypes = [vp, vp, vp, i, i, i, i, i, i, u, f, vp]
    L.usf_corr_fwd_into.restype = i
    L.usf_corr_fwd_into.argtypes = [vp, vp, vp, ctypes.c_longlong, i, i, i, i, i, i, u, f, vp]
    L.usf_corr_bwd.restype = i
    L.usf_corr_bwd.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, i, u, vp]
    L.usf_warp_fwd.restype = i
    L.usf_warp_fwd.argtypes = [vp, vp, vp, i, i, i, i, i, i, u, vp]
    L.usf_warp_bwd_workspace_bytes.restype = sz
    L.usf_warp_bwd_workspace_bytes.argtypes = [i, i, i, i, u]
    L.usf_warp_bwd_workspace_clean_bytes.restype = sz
    L.usf_warp_bwd_workspace_clean_bytes.argtypes = [i, i, i, i, u]
    L.usf_warp_bwd.restype = i
    L.usf_warp_bwd.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, i, u, vp, sz, vp]
    L.usf_occ_bidirection.restype = i
    L.usf_occ_bidirection.argtypes = [vp, vp, vp, i, i, i, f, f, u, vp]
    L.usf_occ_backward_workspace_bytes.restype = sz
    L.usf_occ_backward_workspace_bytes.argtypes = [i, i, i]
    L.usf_occ_backward.restype = i
    L.usf_occ_backward.argtypes = [vp, vp, vp, i, i, i, f, vp, sz, vp]
    L.usf_segpool_max_segments.restype = i
    L.usf_segpool_fwd.restype = i
    L.usf_segpool_fwd.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i, vp]
    L.usf_segpool_bwd.restype = i
    L.usf_segpool_bwd.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i, vp]
    if L.usf_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI {L.usf_abi_version()} != expected {ABI_VERSION}; rebuild")
    _lib = L
    return L


def check(status, what):
    if status != 0:
        L = load()
        msg = L.usf_status_string(status).decode()
        extra = f" (cudaError {L.usf_last_cuda_error()})" if status == -4 else ""
        raise UsfError(f"{what}: {msg}{extra}")
