#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

Imports /root/reference/models/correlation_native.py and
/root/reference/utils/warp_utils.py (namespace packages, nothing copied), runs
them on seeded CPU fp32 inputs through autograd, and stores inputs + outputs.
The reference has no golden vectors of its own for this path, so
these fixtures are what pins oracle/ and, through it, the CUDA kernels.

Usage:  python oracle/make_golden.py [--ref /root/reference]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")

# name, (B, C, H, W)
CORR_CASES = [
    ("tiny", (2, 5, 7, 9)),
    ("smaller_than_window", (1, 7, 3, 5)),
    ("kitti_l0", (1, 192, 4, 13)),
    ("kitti_l1", (1, 128, 8, 26)),
    ("c32_tile_edges", (1, 32, 12, 40)),
    ("c196_ragged", (1, 196, 5, 11)),
    ("one_row", (2, 8, 1, 6)),
]

# name, (B, C, H, W), pad, flow kind
WARP_CASES = [
    ("border_small", (2, 3, 9, 11), "border", "gauss"),
    ("zeros_small", (2, 3, 9, 11), "zeros", "gauss"),
    ("border_wild", (1, 4, 8, 26), "border", "wild"),
    ("zeros_wild", (1, 4, 8, 26), "zeros", "wild"),
    ("border_integer", (1, 5, 6, 13), "border", "integer"),
    ("zeros_integer", (1, 5, 6, 13), "zeros", "integer"),
    ("border_zero_flow", (1, 2, 4, 13), "border", "zero"),
    ("border_c32", (1, 32, 16, 52), "border", "gauss"),
    ("zeros_c2_flowlike", (2, 2, 16, 20), "zeros", "gauss"),
    ("border_img_wide", (1, 3, 32, 104), "border", "gauss"),
]

# name, (B, H, W), flow kind -- the two occlusion masks of utils/warp_utils.py:109-126
OCC_CASES = [
    ("smooth", (2, 16, 52), "smooth"),
    ("gauss", (2, 9, 11), "gauss"),
    ("wild", (1, 8, 26), "wild"),
    ("integer", (1, 6, 13), "integer"),
    ("zero", (1, 4, 13), "zero"),
    ("collapse", (1, 7, 9), "collapse"),
    ("kitti_s2", (1, 64, 208), "smooth"),
]


def make_flow(kind, B, H, W, gen):
    if kind == "smooth":  # a few pixels of low-frequency motion, like a real flow field
        coarse = 4.0 * torch.randn(B, 2, max(H // 8, 2), max(W // 8, 2), generator=gen)
        return torch.nn.functional.interpolate(coarse, (H, W), mode="bilinear", align_corners=True)
    if kind == "collapse":  # every pixel lands on (2.25, 3.5): thousands of taps on four targets
        ys, xs = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
        return torch.stack([2.25 - xs, 3.5 - ys]).expand(B, 2, H, W).clone()
    if kind == "gauss":
        return 3.0 * torch.randn(B, 2, H, W, generator=gen)
    if kind == "wild":
        f = torch.randn(B, 2, H, W, generator=gen)
        f[:, 0] *= W / 4.0
        f[:, 1] *= H / 4.0
        return f
    if kind == "integer":
        return torch.randint(-3, 4, (B, 2, H, W), generator=gen).float()
    if kind == "zero":
        return torch.zeros(B, 2, H, W)
    raise ValueError(kind)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--only", choices=["occ"], help="write only this fixture file")
    args = ap.parse_args()
    sys.path.insert(0, args.ref)
    from models.correlation_native import Correlation  # noqa: E402  (reference, unmodified)
    from utils.warp_utils import flow_warp  # noqa: E402  (reference, unmodified)
    from utils import warp_utils as ref_warp_utils  # noqa: E402

    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    meta = {"torch": torch.__version__}

    blobs = {}
    for i, (name, (B, H, W), kind) in enumerate(OCC_CASES):
        gen = torch.Generator().manual_seed(9876 + i)
        flow12 = make_flow(kind, B, H, W, gen)
        # a backward flow that is roughly consistent with flow12, so that the masks are mixed
        flow21 = -flow12
        for _ in range(10):  # fixed point of flow21(y) = -flow12(y + flow21(y))
            flow21 = -flow_warp(flow12, flow21, pad="border")
        flow21 = flow21 + 0.45 * torch.randn(B, 2, H, W, generator=gen)
        if kind == "collapse":  # the splat is driven by flow21: collapse that one
            flow21 = flow12.clone()
        base = ref_warp_utils.mesh_grid(B, H, W).type_as(flow21)
        for k, v in (("flow12", flow12), ("flow21", flow21),
                     ("occ_bidirection", ref_warp_utils.get_occu_mask_bidirection(flow12, flow21)),
                     ("occ_backward", ref_warp_utils.get_occu_mask_backward(flow21)),
                     ("range_map", ref_warp_utils.get_corresponding_map(base + flow21))):
            blobs[f"{name}.{k}"] = v.detach().numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "occ_golden.npz"), **blobs)
    print("occ_golden.npz", sum(v.nbytes for v in blobs.values()) // 1024, "KiB raw; occluded fractions:",
          {k: round(float(v.mean()), 3) for k, v in blobs.items() if ".occ_" in k})
    if args.only == "occ":
        return

    corr = Correlation(max_displacement=4, kernel_size=1, stride1=1, stride2=1, corr_multiply=1)
    blobs = {}
    for i, (name, (B, C, H, W)) in enumerate(CORR_CASES):
        gen = torch.Generator().manual_seed(1234 + i)
        x1 = torch.randn(B, C, H, W, generator=gen).requires_grad_(True)
        x2 = torch.randn(B, C, H, W, generator=gen).requires_grad_(True)
        gout = torch.randn(B, 81, H, W, generator=gen)
        out = corr(x1, x2)
        out.backward(gout)
        for k, v in (("x1", x1), ("x2", x2), ("gout", gout), ("out", out), ("gx1", x1.grad),
                     ("gx2", x2.grad)):
            blobs[f"{name}.{k}"] = v.detach().numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "corr_golden.npz"), **blobs)
    print("corr_golden.npz", sum(v.nbytes for v in blobs.values()) // 1024, "KiB raw")

    blobs = {}
    for i, (name, (B, C, H, W), pad, kind) in enumerate(WARP_CASES):
        gen = torch.Generator().manual_seed(4321 + i)
        x = torch.randn(B, C, H, W, generator=gen).requires_grad_(True)
        flow = make_flow(kind, B, H, W, gen).requires_grad_(True)
        gout = torch.randn(B, C, H, W, generator=gen)
        out = flow_warp(x, flow, pad=pad)
        out.backward(gout)
        for k, v in (("x", x), ("flow", flow), ("gout", gout), ("out", out), ("gx", x.grad),
                     ("gflow", flow.grad)):
            blobs[f"{name}.{k}"] = v.detach().numpy().astype(np.float32)
        blobs[f"{name}.pad"] = np.array(pad)
    np.savez_compressed(os.path.join(OUT, "warp_golden.npz"), **blobs)
    print("warp_golden.npz", sum(v.nbytes for v in blobs.values() if v.dtype == np.float32) // 1024,
          "KiB raw")
    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write(
            "Golden vectors for the correlation / flow_warp hot path and the two occlusion masks.\n\n"
            "Generated by `python oracle/make_golden.py` in the build container from the UNMODIFIED\n"
            "reference (`/root/reference/models/correlation_native.py`, `utils/warp_utils.py`),\n"
            f"CPU fp32, torch {meta['torch']}, single thread.  Keys are `<case>.<tensor>`; outputs\n"
            "(`out`, `gx*`, `gflow`) come from the reference's forward and autograd backward on the\n"
            "stored inputs.  `occ_golden.npz`: `get_occu_mask_bidirection(flow12, flow21)`,\n"
            "`get_occu_mask_backward(flow21)` and `get_corresponding_map(mesh_grid + flow21)` of\n"
            "`utils/warp_utils.py`.  The reference itself ships no golden vectors for this path.\n")


if __name__ == "__main__":
    main()
