"""Drop-in `flow_warp` (same signature as the reference's utils/warp_utils.py:97-106).

Call sites it serves: models/pwclite.py:302,355 (features, both gradients),
losses/flow_loss.py:130-131 (images, flow gradient only), utils/warp_utils.py:110
(pad="zeros", no gradient) and transforms/ar_transforms/oc_transforms.py:147-148.
The pixel grid, normalisation and bilinear sampling all happen inside one kernel; there is no
CPU mesh_grid and no host-to-device copy per call (warp_utils.py:7-13, :100).
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops


class FlowWarpFunction(Function):
    @staticmethod
    def forward(ctx, x, flow12, pad, reciprocal):
        ctx.save_for_backward(x, flow12)
        ctx.pad = pad
        ctx.reciprocal = reciprocal
        return ops.warp_fwd(x, flow12, pad, reciprocal)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        x, flow12 = ctx.saved_tensors
        gx, gflow = ops.warp_bwd(x, flow12, grad_output, ctx.pad, ctx.needs_input_grad[0],
                                 ctx.needs_input_grad[1], ctx.reciprocal)
        if gflow is not None and gflow.dtype != flow12.dtype:
            gflow = gflow.to(flow12.dtype)
        return gx, gflow, None, None


def flow_warp(x, flow12, pad="border", mode="bilinear", coord_reciprocal=False):
    """x: [B,C,H,W], flow12: [B,2,H,W] (channel 0 along W, channel 1 along H) -> [B,C,H,W].

    coord_reciprocal=True reproduces ATen's CUDA `tensor / scalar` (multiply by the reciprocal) in
    norm_grid; the default reproduces the CPU reference bit-for-bit in the coordinates."""
    if mode != "bilinear":
        raise NotImplementedError(f"flow_warp(mode={mode!r}): every reference call site uses 'bilinear'")
    if pad not in ("border", "zeros"):
        raise ValueError(f"pad={pad!r}: expected 'border' or 'zeros'")
    return FlowWarpFunction.apply(x, flow12, pad, bool(coord_reciprocal))
