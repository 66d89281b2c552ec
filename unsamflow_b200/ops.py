"""Raw (no autograd) tensor front end of the C ABI: validation + pointer extraction only.

Mirrors what the reference's C++ host layer does (models/correlation_package/
correlation_cuda.cc:10-86, 88-165: shape math, output allocation, stream lookup), minus the
scratch tensors and fills the new kernels do not need.
"""
import torch

from . import _lib

_PADS = {"border": _lib.PAD_BORDER, "zeros": _lib.PAD_ZEROS}
_DTYPES = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}
# bf16 storage / fp32 accumulate.  The C ABI handles bf16 for every shape with the direct kernels; above
# this many pixels the host mirror instead widens to fp32, runs the tiled TMA kernels and narrows the
# result (same arithmetic: bf16 -> fp32 is exact, sums are fp32, one rounding at the end), which is much
# faster until native bf16 tiles land.
BF16_NATIVE_MAX_PIXELS = 4096
MAX_DISPLACEMENT = 4
NUM_DISPLACEMENTS = 81


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _feat(t, name):
    if not isinstance(t, torch.Tensor) or t.dim() != 4:
        raise ValueError(f"{name}: expected a 4-D tensor [B,C,H,W]")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: unsamflow_b200 runs on CUDA tensors only (got {t.device}); "
                           "there is no CPU path")
    if t.dtype not in _DTYPES:
        raise NotImplementedError(f"{name}: dtype {t.dtype} is not supported (float32 or bfloat16 storage)")
    return t.contiguous()


def _same(a, b, na, nb):
    if a.shape != b.shape:
        raise ValueError(f"{na}.shape {tuple(a.shape)} != {nb}.shape {tuple(b.shape)}")
    if a.device != b.device:
        raise RuntimeError(f"{na} and {nb} are on different devices")


def _check_out(t, shape, dtype, device, name):
    """A caller-supplied buffer goes to the kernels as a raw pointer: it must be exactly what the kernel
    writes (shape, dtype, device, contiguous) or the device would silently write out of bounds."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise ValueError(f"{name}=: expected a tensor, got {type(t).__name__}")
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}=: shape {tuple(t.shape)} != {tuple(shape)}")
    if t.dtype != dtype:
        raise ValueError(f"{name}=: dtype {t.dtype} != {dtype}")
    if t.device != device:
        raise ValueError(f"{name}=: device {t.device} != {device}")
    if not t.is_contiguous():
        raise ValueError(f"{name}=: must be contiguous")
    return t


def _check_ws(ws, nbytes, device, name="workspace"):
    if not isinstance(ws, torch.Tensor) or ws.device != device or not ws.is_contiguous():
        raise ValueError(f"{name}=: expected a contiguous tensor on {device}")
    if ws.numel() * ws.element_size() < nbytes:
        raise ValueError(f"{name} has {ws.numel() * ws.element_size()} bytes, need {nbytes}")
    if ws.data_ptr() % 16:
        raise ValueError(f"{name}=: must be 16-byte aligned")
    return ws


def _flags(leaky=False, reciprocal=False, atomic=False, ws_clean=False):
    return ((_lib.FLAG_LEAKY_RELU if leaky else 0) | (_lib.FLAG_COORD_RECIPROCAL if reciprocal else 0)
            | (_lib.FLAG_WARP_BWD_ATOMIC if atomic else 0) | (_lib.FLAG_WORKSPACE_CLEAN if ws_clean else 0))


def _is_channel_slice(t, shape, dtype, device):
    """out= may be a view of 81 consecutive channels of a wider contiguous [B,Ctot,H,W] tensor (the estimator's
    concat buffer): dense inside an image, a larger batch stride."""
    B, K, H, W = shape
    return (isinstance(t, torch.Tensor) and tuple(t.shape) == tuple(shape) and t.dtype == dtype and t.device == device
            and t.stride()[1:] == (H * W, W, 1) and (B == 1 or t.stride(0) >= K * H * W))


def corr_fwd(x1, x2, leaky_slope=None, out=None):
    """[B,C,H,W] x2 -> [B,81,H,W]; optional fused LeakyReLU(leaky_slope).  out= is a preallocated contiguous cost
    volume, or a channel slice `buf[:, k:k+81]` of a wider contiguous tensor (usf_corr_fwd_into)."""
    L = _lib.load()
    x1, x2 = _feat(x1, "x1"), _feat(x2, "x2")
    _same(x1, x2, "x1", "x2")
    if x1.dtype != x2.dtype:
        raise ValueError(f"x1 is {x1.dtype} but x2 is {x2.dtype}")
    B, C, H, W = x1.shape
    shape = (B, NUM_DISPLACEMENTS, H, W)
    sliced = out is not None and not out.is_contiguous() and _is_channel_slice(out, shape, x1.dtype, x1.device)
    if not sliced:
        _check_out(out, shape, x1.dtype, x1.device, "out")
    if x1.numel() == 0:  # empty batch: the reference returns an empty cost volume too
        return x1.new_empty(shape) if out is None else out
    if x1.dtype == torch.bfloat16 and B * H * W > BF16_NATIVE_MAX_PIXELS and W % 8 != 0:
        # (W % 8 == 0 takes the native bf16 tiled TMA kernel, csrc/corr_fwd_bf16.cu)
        wide = corr_fwd(x1.float(), x2.float(), leaky_slope)
        if out is None:
            return wide.to(torch.bfloat16)
        out.copy_(wide)
        return out
    if out is None:
        out = torch.empty(shape, dtype=x1.dtype, device=x1.device)
    flags = _flags(leaky=leaky_slope is not None)
    with torch.cuda.device(x1.device):
        if sliced:
            st = L.usf_corr_fwd_into(x1.data_ptr(), x2.data_ptr(), out.data_ptr(), out.stride(0) if B > 1 else
                                     NUM_DISPLACEMENTS * H * W, B, C, H, W, MAX_DISPLACEMENT, _DTYPES[x1.dtype], flags,
                                     float(leaky_slope or 0.0), _stream(x1))
        else:
            st = L.usf_corr_fwd(x1.data_ptr(), x2.data_ptr(), out.data_ptr(), B, C, H, W, MAX_DISPLACEMENT,
                                _DTYPES[x1.dtype], flags, float(leaky_slope or 0.0), _stream(x1))
    _lib.check(st, "usf_corr_fwd_into" if sliced else "usf_corr_fwd")
    return out


def corr_bwd(x1, x2, gout, need_gx1=True, need_gx2=True, gx1=None, gx2=None):
    """-> (gx1, gx2); pass preallocated gx1= / gx2= to avoid allocations (CUDA-graph friendly)."""
    L = _lib.load()
    x1, x2, gout = _feat(x1, "x1"), _feat(x2, "x2"), _feat(gout, "grad_output")
    _same(x1, x2, "x1", "x2")
    B, C, H, W = x1.shape
    if tuple(gout.shape) != (B, NUM_DISPLACEMENTS, H, W):
        raise ValueError(f"grad_output.shape {tuple(gout.shape)} != {(B, NUM_DISPLACEMENTS, H, W)}")
    if not (x1.dtype == x2.dtype == gout.dtype):
        raise ValueError(f"dtype mismatch: {x1.dtype}, {x2.dtype}, {gout.dtype}")
    if gout.device != x1.device:
        raise RuntimeError("x1 and grad_output are on different devices")
    _check_out(gx1, x1.shape, x1.dtype, x1.device, "gx1")
    _check_out(gx2, x2.shape, x2.dtype, x2.device, "gx2")
    if x1.numel() == 0:
        return (torch.empty_like(x1) if need_gx1 else None, torch.empty_like(x2) if need_gx2 else None)
    if x1.dtype == torch.bfloat16 and B * H * W > BF16_NATIVE_MAX_PIXELS and W % 8 != 0:
        # (W % 8 == 0 takes the native bf16 tiled TMA kernel, csrc/corr_bwd_bf16.cu)
        w1, w2 = corr_bwd(x1.float(), x2.float(), gout.float(), need_gx1, need_gx2)
        res = []
        for wide, dst, need in ((w1, gx1, need_gx1), (w2, gx2, need_gx2)):
            if not need:
                res.append(None)
            elif dst is None:
                res.append(wide.to(torch.bfloat16))
            else:
                dst.copy_(wide)
                res.append(dst)
        return tuple(res)
    if need_gx1 and gx1 is None:
        gx1 = torch.empty_like(x1)
    if need_gx2 and gx2 is None:
        gx2 = torch.empty_like(x2)
    with torch.cuda.device(x1.device):
        st = L.usf_corr_bwd(x1.data_ptr(), x2.data_ptr(), gout.data_ptr(),
                            gx1.data_ptr() if need_gx1 else None, gx2.data_ptr() if need_gx2 else None,
                            B, C, H, W, MAX_DISPLACEMENT, _DTYPES[x1.dtype], 0, _stream(x1))
    _lib.check(st, "usf_corr_bwd")
    return gx1, gx2


def _flow(flow, x):
    if not isinstance(flow, torch.Tensor) or flow.dim() != 4:
        raise ValueError("flow12: expected a 4-D tensor [B,2,H,W]")
    B, _, H, W = x.shape
    if tuple(flow.shape) != (B, 2, H, W):
        raise ValueError(f"flow12.shape {tuple(flow.shape)} != {(B, 2, H, W)}")
    if flow.device != x.device:
        raise RuntimeError("x and flow12 are on different devices")
    return flow.float().contiguous()


def _pad(pad):
    try:
        return _PADS[pad]
    except KeyError:
        raise ValueError(f"pad={pad!r}: expected 'border' or 'zeros'") from None


def warp_fwd(x, flow, pad="border", reciprocal=False, out=None):
    L = _lib.load()
    x = _feat(x, "x")
    flow = _flow(flow, x)
    B, C, H, W = x.shape
    if H < 2 or W < 2:
        raise ValueError("flow_warp needs H, W >= 2 (the reference divides by size-1)")
    _check_out(out, x.shape, x.dtype, x.device, "out")
    if out is None:
        out = torch.empty_like(x)
    if x.numel() == 0:
        return out
    with torch.cuda.device(x.device):
        st = L.usf_warp_fwd(x.data_ptr(), flow.data_ptr(), out.data_ptr(), B, C, H, W, _pad(pad), _DTYPES[x.dtype],
                            _flags(reciprocal=reciprocal), _stream(x))
    _lib.check(st, "usf_warp_fwd")
    return out


def warp_bwd_workspace_bytes(x, atomic=False):
    B, C, H, W = x.shape
    return int(_lib.load().usf_warp_bwd_workspace_bytes(B, C, H, W, _flags(atomic=atomic)))


def warp_bwd_workspace(x, atomic=False):
    """A persistent workspace for warp_bwd(..., workspace=ws, workspace_clean=True): allocated once per call site
    with its counter region zeroed; every warp_bwd call leaves it zeroed again (usf_warp_bwd_workspace_clean_bytes),
    so the per-call memset disappears from the step.  One workspace serves one call at a time."""
    B, C, H, W = x.shape
    L = _lib.load()
    flags = _flags(atomic=atomic)
    ws = torch.empty(int(L.usf_warp_bwd_workspace_bytes(B, C, H, W, flags)), dtype=torch.uint8, device=x.device)
    ws[:int(L.usf_warp_bwd_workspace_clean_bytes(B, C, H, W, flags))].zero_()
    return ws


def warp_bwd(x, flow, gout, pad="border", need_gx=True, need_gflow=True, reciprocal=False, atomic=False,
             gx=None, gflow=None, workspace=None, workspace_clean=False):
    """-> (gx, gflow); gx= / gflow= / workspace= (uint8 tensor) may be preallocated.  workspace_clean=True:
    `workspace` came from warp_bwd_workspace() (counters zero on entry, left zero on exit)."""
    L = _lib.load()
    x, gout = _feat(x, "x"), _feat(gout, "grad_output")
    _same(x, gout, "x", "grad_output")
    if x.dtype != gout.dtype:
        raise ValueError(f"x is {x.dtype} but grad_output is {gout.dtype}")
    flow = _flow(flow, x)
    B, C, H, W = x.shape
    _check_out(gx, x.shape, x.dtype, x.device, "gx")
    _check_out(gflow, flow.shape, torch.float32, x.device, "gflow")
    if x.numel() == 0:
        return (torch.empty_like(x) if need_gx else None, torch.empty_like(flow) if need_gflow else None)
    if workspace_clean and workspace is None:
        raise ValueError("workspace_clean=True needs the persistent workspace= it refers to")
    flags = _flags(reciprocal=reciprocal, atomic=atomic, ws_clean=workspace_clean)
    if need_gx and gx is None:
        gx = torch.empty_like(x)
    if need_gflow and gflow is None:
        gflow = torch.empty_like(flow)
    ws, ws_bytes = None, 0
    if need_gx:
        ws_bytes = L.usf_warp_bwd_workspace_bytes(B, C, H, W, flags)
        if ws_bytes:
            ws = workspace if workspace is not None else torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
            _check_ws(ws, ws_bytes, x.device)
    with torch.cuda.device(x.device):
        st = L.usf_warp_bwd(x.data_ptr(), flow.data_ptr(), gout.data_ptr(),
                            gx.data_ptr() if need_gx else None, gflow.data_ptr() if need_gflow else None,
                            B, C, H, W, _pad(pad), _DTYPES[x.dtype], flags,
                            ws.data_ptr() if ws is not None else None, ws_bytes, _stream(x))
    _lib.check(st, "usf_warp_bwd")
    return gx, gflow


def _flow_pair(t, name):
    if not isinstance(t, torch.Tensor) or t.dim() != 4 or t.shape[1] != 2:
        raise ValueError(f"{name}: expected a 4-D flow tensor [B,2,H,W]")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: unsamflow_b200 runs on CUDA tensors only (got {t.device}); "
                           "there is no CPU path")
    return t.detach().float().contiguous()


def occ_bidirection(flow12, flow21, scale=0.01, bias=0.5, reciprocal=False, out=None):
    """[B,2,H,W] x2 -> [B,1,H,W] fp32 mask (1 = occluded); utils/warp_utils.py:109-117 in one kernel."""
    L = _lib.load()
    flow12, flow21 = _flow_pair(flow12, "flow12"), _flow_pair(flow21, "flow21")
    _same(flow12, flow21, "flow12", "flow21")
    B, _, H, W = flow12.shape
    if H < 2 or W < 2:
        raise ValueError("get_occu_mask_bidirection needs H, W >= 2 (flow_warp divides by size-1)")
    _check_out(out, (B, 1, H, W), torch.float32, flow12.device, "out")
    if out is None:
        out = torch.empty((B, 1, H, W), dtype=torch.float32, device=flow12.device)
    if flow12.numel() == 0:
        return out
    with torch.cuda.device(flow12.device):
        st = L.usf_occ_bidirection(flow12.data_ptr(), flow21.data_ptr(), out.data_ptr(), B, H, W, float(scale),
                                   float(bias), _flags(reciprocal=reciprocal), _stream(flow12))
    _lib.check(st, "usf_occ_bidirection")
    return out


def occ_backward_workspace_bytes(flow21):
    B, _, H, W = flow21.shape
    return int(_lib.load().usf_occ_backward_workspace_bytes(B, H, W))


def occ_backward(flow21, th=0.2, want_range_map=False, out=None, workspace=None):
    """[B,2,H,W] -> [B,1,H,W] fp32 mask (1 = no source pixel maps here); utils/warp_utils.py:120-126.
    want_range_map=True returns (mask, range map) -- get_corresponding_map's value (:26-94)."""
    L = _lib.load()
    flow21 = _flow_pair(flow21, "flow21")
    B, _, H, W = flow21.shape
    _check_out(out, (B, 1, H, W), torch.float32, flow21.device, "out")
    if out is None:
        out = torch.empty((B, 1, H, W), dtype=torch.float32, device=flow21.device)
    rng = torch.empty_like(out) if want_range_map else None
    if flow21.numel() == 0:
        return (out, rng) if want_range_map else out
    ws_bytes = L.usf_occ_backward_workspace_bytes(B, H, W)
    ws = workspace if workspace is not None else torch.empty(ws_bytes, dtype=torch.uint8, device=flow21.device)
    _check_ws(ws, ws_bytes, flow21.device)
    with torch.cuda.device(flow21.device):
        st = L.usf_occ_backward(flow21.data_ptr(), out.data_ptr(), rng.data_ptr() if want_range_map else None,
                                B, H, W, float(th), ws.data_ptr(), ws_bytes, _stream(flow21))
    _lib.check(st, "usf_occ_backward")
    return (out, rng) if want_range_map else out


def launch_count():
    return int(_lib.load().usf_launch_count())


def reset_launch_count():
    _lib.load().usf_reset_launch_count()
