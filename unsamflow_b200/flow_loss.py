"""Unsupervised photometric loss around the hot path: the OTHER caller of flow_warp.

Restates the stage-0 configuration of losses/flow_loss.py:83-259 (configs/kitti_base.json:27-40:
w_l1 0.15, w_ssim 0.85, w_ternary 0, w_sm 0, occ_from_back, w_ph_scales [1,1,1,1,0], warp_pad border)
so that a full training step can run on the GPU box without the reference tree.  The two call sites
this repo accelerates are

    im1_recons = flow_warp(im2_scaled, flow[:, :2], pad=warp_pad)     losses/flow_loss.py:130
    im2_recons = flow_warp(im1_scaled, flow[:, 2:], pad=warp_pad)     losses/flow_loss.py:131

and, with occ_from_back, the two range-map occlusion masks

    occu_mask1 = 1 - get_occu_mask_backward(flow[:, 2:], th=0.2)      losses/flow_loss.py:101
    occu_mask2 = 1 - get_occu_mask_backward(flow[:, :2], th=0.2)      losses/flow_loss.py:102

(unsamflow_b200/occlusion.py).  Everything else (area resize, SSIM) is elementwise PyTorch as in the
reference and out of scope.
"""
import torch
import torch.nn.functional as F
from torch import nn

from .occlusion import get_occu_mask_backward, get_occu_mask_bidirection
from .warp import flow_warp


def ssim_distance(x, y):
    """3x3 SSIM dissimilarity, losses/loss_blocks.py:53-72."""
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    pool = lambda t: F.avg_pool2d(t, 3, 1, 0)  # noqa: E731
    mu_x, mu_y = pool(x), pool(y)
    sig_x = pool(x * x) - mu_x * mu_x
    sig_y = pool(y * y) - mu_y * mu_y
    sig_xy = pool(x * y) - mu_x * mu_y
    ssim = ((2 * mu_x * mu_y + c1) * (2 * sig_xy + c2)) / ((mu_x * mu_x + mu_y * mu_y + c1) * (sig_x + sig_y + c2))
    return torch.clamp((1 - ssim) / 2, 0, 1)


class UnFlowLoss(nn.Module):
    def __init__(self, w_l1=0.15, w_ssim=0.85, w_scales=(1.0, 1.0, 1.0, 1.0, 0.0), warp_pad="border",
                 occ_from_back=True):
        super().__init__()
        self.w_l1, self.w_ssim, self.w_scales, self.warp_pad = w_l1, w_ssim, w_scales, warp_pad
        self.occ_from_back = occ_from_back

    def photometric(self, target, recons, vis):
        terms = []
        if self.w_l1 > 0:
            terms.append(self.w_l1 * (target - recons).abs() * vis)
        if self.w_ssim > 0:
            terms.append(self.w_ssim * ssim_distance(recons * vis, target * vis))
        return sum(t.mean() for t in terms) / (vis.mean() + 1e-6)

    def forward(self, pyramid_flows, img1, img2):
        """pyramid_flows: list of [B,4,h,w] = cat(flow12, flow21) per scale, finest first
        (trainer/kitti_trainer_ar.py:115-127)."""
        top = pyramid_flows[0]
        # occ_from_back (losses/flow_loss.py:99-106): range-map masks of the finest flows, no gradient
        if self.occ_from_back:
            vis1 = 1 - get_occu_mask_backward(top[:, 2:].detach(), th=0.2)
            vis2 = 1 - get_occu_mask_backward(top[:, :2].detach(), th=0.2)
        else:  # forward-backward consistency (losses/flow_loss.py:104-106)
            vis1 = 1 - get_occu_mask_bidirection(top[:, :2].detach(), top[:, 2:].detach())
            vis2 = 1 - get_occu_mask_bidirection(top[:, 2:].detach(), top[:, :2].detach())
        total = 0.0
        for i, flow in enumerate(pyramid_flows):
            if self.w_scales[i] <= 0:
                continue
            h, w = flow.shape[-2:]
            im1 = F.interpolate(img1, (h, w), mode="area")
            im2 = F.interpolate(img2, (h, w), mode="area")
            rec1 = flow_warp(im2, flow[:, :2], pad=self.warp_pad)
            rec2 = flow_warp(im1, flow[:, 2:], pad=self.warp_pad)
            v1 = vis1 if i == 0 else F.interpolate(vis1, (h, w), mode="nearest")
            v2 = vis2 if i == 0 else F.interpolate(vis2, (h, w), mode="nearest")
            # with_bk: the two directions are averaged (losses/flow_loss.py:143-148)
            total = total + self.w_scales[i] * 0.5 * (self.photometric(im1, rec1, v1) + self.photometric(im2, rec2, v2))
        return total
