"""Drop-in occlusion masks (same names and signatures as the reference's utils/warp_utils.py:109-126).

    get_occu_mask_bidirection(flow12, flow21, scale=0.01, bias=0.5)   losses/flow_loss.py:108-113
    get_occu_mask_backward(flow21, th=0.2)                            losses/flow_loss.py:100-106

Both return a float mask [B,1,H,W] (1 = occluded) of the flow's dtype and carry no gradient: in the
reference the result goes through a comparison, and the loss detaches it.  One kernel (bidirection) or
memset + two kernels (backward) instead of a CPU mesh_grid, an H2D copy and 20-40 ATen kernels; the
backward mask is bit-identical from run to run (integer accumulation), which the reference's float-atomic
scatter_add_ on a GPU is not.  CUDA tensors only; there is no CPU path.
"""
import torch

from . import ops


@torch.no_grad()
def get_occu_mask_bidirection(flow12, flow21, scale=0.01, bias=0.5):
    return ops.occ_bidirection(flow12, flow21, scale, bias).to(flow12.dtype)


@torch.no_grad()
def get_occu_mask_backward(flow21, th=0.2):
    return ops.occ_backward(flow21, th).to(flow21.dtype)


@torch.no_grad()
def range_map_from_flow(flow21):
    """Range map of `mesh_grid + flow21`: how many source pixels (bilinearly split) land on each target.
    This is what the reference's get_occu_mask_backward hands to get_corresponding_map
    (utils/warp_utils.py:122-124), but the argument here is the FLOW -- the pixel grid is added inside
    the kernel -- which is why it does not carry the reference's name (a caller ported from the reference
    would otherwise add the grid twice).  Recovering the flow from absolute coordinates by subtraction is
    not exact in fp32, so no coordinate-taking variant is offered."""
    return ops.occ_backward(flow21, 0.2, want_range_map=True)[1].to(flow21.dtype)
