"""Segment max-pooling of the "+mf" mask-feature branch (models/pwclite.py:317-340) on the usf kernels.

    segment_max_pool(feat, seg, num_segments=None) -> pooled features, same shape as `feat`

`feat` [B,C,h,w] (float32 / bfloat16, CUDA), `seg` [B,1,h,w] or [B,h,w] holding integer segment ids (any dtype; the
reference gets them as floats from F.interpolate(mode="nearest") and calls .long()).  Every pixel receives the
maximum of `feat` over its segment, clamped at 0 from below unless the segment covers the whole image -- exactly what
the reference's `one_hot x feature -> amax -> one_hot-weighted sum` produces, without its [B,C,h,w,N] temporaries
(0.5-3.5 GB at B=8, 64x208).  Differentiable wrt `feat`; bit-reproducible forward and backward.
`num_segments`: number of one-hot classes (max id + 1).  None reads it from `seg` (one device->host sync per call,
as the reference's one_hot does); pass it (e.g. 256 for uint8 SAM id maps) to keep the step sync-free / capturable.
There is no CPU path.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib, ops


def segpool_fwd(feat, seg, num_segments, want_argmax=True, out=None):
    L = _lib.load()
    feat = ops._feat(feat, "feat")
    B, C, H, W = feat.shape
    if not isinstance(seg, torch.Tensor) or seg.device != feat.device or tuple(seg.shape) != (B, H, W) \
            or seg.dtype != torch.int32 or not seg.is_contiguous():
        raise ValueError(f"seg: expected a contiguous int32 tensor {(B, H, W)} on {feat.device}")
    if not 1 <= num_segments <= L.usf_segpool_max_segments():
        raise ValueError(f"num_segments={num_segments}: supported range is 1..{L.usf_segpool_max_segments()}")
    ops._check_out(out, feat.shape, feat.dtype, feat.device, "out")
    if out is None:
        out = torch.empty_like(feat)
    arg = torch.empty((B, C, num_segments), dtype=torch.int32, device=feat.device) if want_argmax else None
    if feat.numel() == 0:
        return out, arg
    with torch.cuda.device(feat.device):
        st = L.usf_segpool_fwd(feat.data_ptr(), seg.data_ptr(), out.data_ptr(), arg.data_ptr() if want_argmax else None,
                               B, C, H, W, int(num_segments), ops._DTYPES[feat.dtype], ops._stream(feat))
    _lib.check(st, "usf_segpool_fwd")
    return out, arg


def segpool_bwd(gout, seg, argmax, num_segments):
    L = _lib.load()
    gout = ops._feat(gout, "grad_output")
    B, C, H, W = gout.shape
    gfeat = torch.empty_like(gout)
    if gout.numel() == 0:
        return gfeat
    with torch.cuda.device(gout.device):
        st = L.usf_segpool_bwd(gout.data_ptr(), seg.data_ptr(), argmax.data_ptr(), gfeat.data_ptr(), B, C, H, W,
                               int(num_segments), ops._DTYPES[gout.dtype], ops._stream(gout))
    _lib.check(st, "usf_segpool_bwd")
    return gfeat


class SegmentMaxPoolFunction(Function):
    @staticmethod
    def forward(ctx, feat, seg, num_segments):
        out, arg = segpool_fwd(feat, seg, num_segments, want_argmax=feat.requires_grad)
        ctx.num_segments = num_segments
        if arg is not None:
            ctx.save_for_backward(seg, arg)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        seg, arg = ctx.saved_tensors
        return segpool_bwd(grad_output, seg, arg, ctx.num_segments), None, None


def segment_max_pool(feat, seg, num_segments=None):
    if not isinstance(feat, torch.Tensor) or feat.dim() != 4:
        raise ValueError("feat: expected a 4-D tensor [B,C,h,w]")
    if not feat.is_cuda:
        raise RuntimeError(f"feat: unsamflow_b200 runs on CUDA tensors only (got {feat.device}); there is no CPU path")
    b, _, h, w = feat.shape
    if seg.dim() == 4:
        if seg.shape[1] != 1:
            raise ValueError("seg: expected [B,1,h,w] or [B,h,w]")
        seg = seg[:, 0]
    if tuple(seg.shape) != (b, h, w):
        raise ValueError(f"seg.shape {tuple(seg.shape)} does not match feat {(b, h, w)}")
    ids = seg.to(torch.int32).contiguous()   # (.long() in the reference: truncation toward zero, same here)
    if num_segments is None:
        num_segments = int(ids.max().item()) + 1 if ids.numel() else 1
    return SegmentMaxPoolFunction.apply(feat, ids, int(num_segments))
