#!/usr/bin/env python
"""Occlusion-mask kernels against the stock-PyTorch formulation of the same functions on the same GPU.

usf: unsamflow_b200.ops.occ_backward / occ_bidirection (preallocated outputs and workspace, CUDA-event timed).
stock: the reference's op sequence (utils/warp_utils.py:26-126) written with plain torch ops, including its
CPU-built pixel grid and H2D copy per call.  Algorithmic bytes per pixel: backward 8 (flow) + 4 (mask) = 12,
bidirection 8 + 8 + 4 = 20.  Prints one JSON line.
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchmarks.train_compare import stock_occ_backward, stock_warp  # noqa: E402
from unsamflow_b200 import ops  # noqa: E402


@torch.no_grad()
def stock_occ_bidirection(flow12, flow21, scale=0.01, bias=0.5):
    back = stock_warp(flow21, flow12, pad="zeros")
    diff = flow12 + back
    mag = (flow12 * flow12).sum(1, keepdim=True) + (back * back).sum(1, keepdim=True)
    return ((diff * diff).sum(1, keepdim=True) > scale * mag + bias).float()


def timed(fn, iters, flush):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(iters):
        if flush is not None:
            flush.add_(1.0)  # 256 MB > L2: inputs come from HBM
        a.record()
        fn()
        b.record()
        b.synchronize()
        total += a.elapsed_time(b)
    return total / iters * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--height", type=int, default=256)
    ap.add_argument("--width", type=int, default=832)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--skip-stock", action="store_true", help="time only the usf kernels")
    args = ap.parse_args()
    B, H, W = args.batch, args.height, args.width
    coarse = 6.0 * torch.randn(B, 2, H // 16, W // 16, device="cuda")
    flow12 = F.interpolate(coarse, (H, W), mode="bilinear", align_corners=True).contiguous()
    flow21 = (-flow12 + 0.5 * torch.randn_like(flow12)).contiguous()
    out = torch.empty(B, 1, H, W, device="cuda")
    ws = torch.empty(ops.occ_backward_workspace_bytes(flow21), dtype=torch.uint8, device="cuda")
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")
    n = B * H * W
    res = {"config": vars(args), "gpu": torch.cuda.get_device_name(0),
           "splat_lane_merge": os.environ.get("USF_OCC_MERGE", "1") != "0",
           "splat_pixels_per_thread": os.environ.get("USF_OCC_ITER", "default"),
           "splat_tile_window": os.environ.get("USF_OCC_TILE", "0") == "1"}
    for name, usf_fn, stock_fn, bytes_px in (
        ("occ_backward", lambda: ops.occ_backward(flow21, out=out, workspace=ws), lambda: stock_occ_backward(flow21), 12),
        ("occ_bidirection", lambda: ops.occ_bidirection(flow12, flow21, out=out), lambda: stock_occ_bidirection(flow12, flow21), 20),
    ):
        ops.reset_launch_count()
        usf_fn()
        launches = ops.launch_count()
        warm, cold = timed(usf_fn, args.iters, None), timed(usf_fn, args.iters, flush)
        stock = float("nan") if args.skip_stock else timed(stock_fn, max(args.iters // 5, 5), None)
        res[name] = {"usf_us_warm": round(warm, 2), "usf_us_cold_l2": round(cold, 2), "stock_pytorch_us": round(stock, 1),
                     "speedup_warm": round(stock / warm, 1), "kernels": launches,
                     "algorithmic_gbs_warm": round(n * bytes_px / warm / 1e3, 1),
                     "algorithmic_gbs_cold_l2": round(n * bytes_px / cold / 1e3, 1),
                     "occluded_fraction": round(float(out.mean()), 4)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
