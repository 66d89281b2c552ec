#!/usr/bin/env python
"""Training step with the usf kernels vs the SAME model with stock PyTorch ops on the same GPU.

The stock arm is what the reference runs when its CUDA package is not built (README.md:17): the
native cost volume (pad + 81 shifted mul/mean + cat) and flow_warp through F.grid_sample -- written
out here (benchmarks/ may not import oracle/).  Prints one JSON line.
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unsamflow_b200.flow_loss as fl  # noqa: E402
import unsamflow_b200.pwclite as pw  # noqa: E402
from unsamflow_b200.train_step import TrainStep  # noqa: E402


class StockCorr(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, x1, x2):
        h, w = x1.shape[-2:]
        x2p = F.pad(x2, [4, 4, 4, 4])
        return torch.cat([(x1 * x2p[:, :, i:i + h, j:j + w]).mean(1, keepdim=True) for i in range(9) for j in range(9)], 1)


def stock_warp(x, flow, pad="border", mode="bilinear"):
    b, _, h, w = x.shape
    xs = torch.arange(0, w).repeat(b, h, 1)
    ys = torch.arange(0, h).repeat(b, w, 1).transpose(1, 2)
    v = torch.stack([xs, ys], 1).type_as(x) + flow          # CPU grid + H2D per call, like the reference
    g = torch.zeros_like(v)
    g[:, 0] = 2.0 * v[:, 0] / (w - 1) - 1.0
    g[:, 1] = 2.0 * v[:, 1] / (h - 1) - 1.0
    return F.grid_sample(x, g.permute(0, 2, 3, 1), mode=mode, padding_mode=pad, align_corners=True)


@torch.no_grad()
def stock_occ_backward(flow21, th=0.2):
    """get_occu_mask_backward with stock PyTorch ops (4-tap scatter_add_ range map; utils/warp_utils.py:26-94,120-126)."""
    b, _, h, w = flow21.shape
    xs = torch.arange(0, w).repeat(b, h, 1)
    ys = torch.arange(0, h).repeat(b, w, 1).transpose(1, 2)
    v = torch.stack([xs, ys], 1).type_as(flow21) + flow21    # CPU grid + H2D per call, like the reference
    x, y = v[:, 0].reshape(b, -1), v[:, 1].reshape(b, -1)
    x0, y0 = torch.floor(x), torch.floor(y)
    acc = torch.zeros(b, h * w, dtype=flow21.dtype, device=flow21.device)
    for dx in (0, 1):
        for dy in (0, 1):
            xi, yi = x0 + dx, y0 + dy
            ok = (xi >= 0) & (xi <= w - 1) & (yi >= 0) & (yi <= h - 1)
            wgt = (1 - (x - xi).abs()) * (1 - (y - yi).abs()) * ok
            acc.scatter_add_(1, (xi.clamp(0, w - 1) + yi.clamp(0, h - 1) * w).long(), wgt)
    return (acc.view(b, 1, h, w).clamp(0, 1) < th).to(flow21.dtype)


def timed(ts, steps):
    for _ in range(3):
        ts.step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        loss = ts.step()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / steps, float(loss)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--height", type=int, default=256)
    ap.add_argument("--width", type=int, default=832)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    res = {}
    for name, bd in (("usf", False), ("usf_batched_directions", True)):
        ms, loss = timed(TrainStep(args.batch, args.height, args.width, "cuda:0", batch_directions=bd), args.steps)
        res[name] = {"ms_per_step": round(ms, 3), "pairs_per_s": round(args.batch / ms * 1e3, 1), "loss": loss}
    pw.Correlation, pw.flow_warp, fl.flow_warp = StockCorr, stock_warp, stock_warp
    fl.get_occu_mask_backward = stock_occ_backward
    ms, loss = timed(TrainStep(args.batch, args.height, args.width, "cuda:0"), args.steps)
    res["stock_pytorch_ops"] = {"ms_per_step": round(ms, 3), "pairs_per_s": round(args.batch / ms * 1e3, 1), "loss": loss}
    res["config"] = vars(args)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
