"""Unsupervised photometric loss around the hot path: the OTHER caller of flow_warp.

Restates the stage-0 configuration of losses/flow_loss.py:83-259 (configs/kitti_base.json:27-40:
w_l1 0.15, w_ssim 0.85, w_ternary 0, w_sm 0, occ_from_back, w_ph_scales [1,1,1,1,0], warp_pad border)
so that a full training step can run on the GPU box without the reference tree.  The two call sites
this repo accelerates are

    im1_recons = flow_warp(im2_scaled, flow[:, :2], pad=warp_pad)     losses/flow_loss.py:130
    im2_recons = flow_warp(im1_scaled, flow[:, 2:], pad=warp_pad)     losses/flow_loss.py:131

Everything else (area resize, SSIM, masks) is elementwise PyTorch as in the reference and out of scope.
"""
import torch
import torch.nn.functional as F
from torch import nn

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


@torch.no_grad()
def occlusion_from_backward_flow(flow21, th=0.2):
    """Forward-splat range map of flow21; pixels nobody maps to are occluded
    (utils/warp_utils.py:26-94, :120-126).  Built on the device (the reference builds the pixel grid on
    the CPU every call)."""
    b, _, h, w = flow21.shape
    ys, xs = torch.meshgrid(torch.arange(h, device=flow21.device, dtype=flow21.dtype),
                            torch.arange(w, device=flow21.device, dtype=flow21.dtype), indexing="ij")
    x = (xs + flow21[:, 0]).reshape(b, -1)
    y = (ys + flow21[:, 1]).reshape(b, -1)
    x0, y0 = torch.floor(x), torch.floor(y)
    acc = torch.zeros(b, h * w, dtype=flow21.dtype, device=flow21.device)
    for dx in (0, 1):
        for dy in (0, 1):
            xi, yi = x0 + dx, y0 + dy
            ok = (xi >= 0) & (xi <= w - 1) & (yi >= 0) & (yi <= h - 1)
            wgt = (1 - (x - xi).abs()) * (1 - (y - yi).abs()) * ok
            idx = (xi.clamp(0, w - 1) + yi.clamp(0, h - 1) * w).long()
            acc.scatter_add_(1, idx, wgt)
    return (acc.view(b, 1, h, w).clamp(0, 1) < th).to(flow21.dtype)


class UnFlowLoss(nn.Module):
    def __init__(self, w_l1=0.15, w_ssim=0.85, w_scales=(1.0, 1.0, 1.0, 1.0, 0.0), warp_pad="border"):
        super().__init__()
        self.w_l1, self.w_ssim, self.w_scales, self.warp_pad = w_l1, w_ssim, w_scales, warp_pad

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
        vis1 = 1 - occlusion_from_backward_flow(top[:, 2:])
        vis2 = 1 - occlusion_from_backward_flow(top[:, :2])
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
