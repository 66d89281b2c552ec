#!/usr/bin/env python
"""Per-op secondary baselines on the SAME GPU: the stock PyTorch formulation of the two ops --
what the reference runs when its CUDA package is not built (native cost volume: pad + 81 shifted mul/mean +
cat; flow_warp: CPU mesh grid + H2D + norm + F.grid_sample) -- against the usf kernels, forward and
forward+backward, KITTI pyramid at B=8 plus the loss-side warps.  Eager calls, CUDA events, median of 20,
warm L2; both sides include their Python launch overhead.  Writes one JSON document.
(benchmarks/ may not import oracle/, so the stock ops are written out here.)
"""
import json
import os
import statistics
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unsamflow_b200 import Correlation, flow_warp  # noqa: E402


def stock_corr(x1, x2):
    h, w = x1.shape[-2:]
    x2p = F.pad(x2, [4, 4, 4, 4])
    return torch.cat([(x1 * x2p[:, :, i:i + h, j:j + w]).mean(1, keepdim=True) for i in range(9) for j in range(9)], 1)


def stock_warp(x, flow, pad="border"):
    b, _, h, w = x.shape
    xs = torch.arange(0, w).repeat(b, h, 1)
    ys = torch.arange(0, h).repeat(b, w, 1).transpose(1, 2)
    v = torch.stack([xs, ys], 1).type_as(x) + flow  # CPU grid + H2D per call, like the reference
    g = torch.zeros_like(v)
    g[:, 0] = 2.0 * v[:, 0] / (w - 1) - 1.0
    g[:, 1] = 2.0 * v[:, 1] / (h - 1) - 1.0
    return F.grid_sample(x, g.permute(0, 2, 3, 1), mode="bilinear", padding_mode=pad, align_corners=True)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return round(statistics.median(ts), 1)


def main():
    g = torch.Generator().manual_seed(0)
    corr = Correlation(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1)
    rows = []
    for C, H, W in [(192, 4, 13), (128, 8, 26), (96, 16, 52), (64, 32, 104), (32, 64, 208)]:
        x1 = torch.randn(8, C, H, W, generator=g).cuda().requires_grad_(True)
        x2 = torch.randn(8, C, H, W, generator=g).cuda().requires_grad_(True)
        go = torch.randn(8, 81, H, W, generator=g).cuda()
        fl = (3.0 * torch.randn(8, 2, H, W, generator=g)).cuda().requires_grad_(True)
        gx = torch.randn(8, C, H, W, generator=g).cuda()
        row = {"shape": [8, C, H, W]}
        for name, cf, wf in (("stock", stock_corr, stock_warp), ("usf", corr, flow_warp)):
            row[f"corr_fwd_{name}_us"] = timed(lambda: cf(x1, x2))
            row[f"corr_fwd_bwd_{name}_us"] = timed(lambda: torch.autograd.grad(cf(x1, x2), (x1, x2), go))
            row[f"warp_fwd_{name}_us"] = timed(lambda: wf(x2, fl))
            row[f"warp_fwd_bwd_{name}_us"] = timed(lambda: torch.autograd.grad(wf(x2, fl), (x2, fl), gx))
        rows.append(row)
        print(json.dumps(row), flush=True)
    for H, W in [(256, 832), (128, 416)]:
        im = torch.rand(8, 3, H, W, generator=g).cuda()
        fl = (3.0 * torch.randn(8, 2, H, W, generator=g)).cuda().requires_grad_(True)
        gi = torch.randn(8, 3, H, W, generator=g).cuda()
        row = {"shape": [8, 3, H, W], "what": "loss-side warp, gradient wrt flow only"}
        for name, wf in (("stock", stock_warp), ("usf", flow_warp)):
            row[f"warp_fwd_{name}_us"] = timed(lambda: wf(im, fl))
            row[f"warp_fwd_bwd_{name}_us"] = timed(lambda: torch.autograd.grad(wf(im, fl), (fl,), gi))
        rows.append(row)
        print(json.dumps(row), flush=True)
    out = {"note": __doc__.strip().split("\n\n")[0], "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0), "rows": rows}
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
