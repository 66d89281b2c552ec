#!/usr/bin/env python
"""Bit-determinism stress: every correlation kernel family (mixed-tile TMA backward, 4-channel thin-grid form,
whole-image kernels, bf16 tiled kernels) and the fused small-image warp backward, 300 runs each under a
second stream that keeps the GPU busy.  A race between a TMA write and a late shared-memory read shows up
here as a handful of mismatching runs (that is how the G-stage release bug of round 1 was found)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unsamflow_b200 import ops  # noqa: E402

RUNS = int(sys.argv[1]) if len(sys.argv) > 1 else 300
g = torch.Generator().manual_seed(7)
side = torch.cuda.Stream()
noise = torch.randn(4096, 4096, device="cuda")
total_bad = 0
for dtype in (torch.float32, torch.bfloat16):
    for shape in [(8, 32, 64, 208), (8, 64, 32, 104), (8, 96, 16, 52), (8, 128, 8, 26), (8, 192, 4, 13), (1, 64, 32, 104)]:
        B, C, H, W = shape
        x1 = torch.randn(shape, generator=g).cuda().to(dtype)
        x2 = torch.randn(shape, generator=g).cuda().to(dtype)
        go = torch.randn(B, 81, H, W, generator=g).cuda().to(dtype)
        r1, r2 = (t.clone() for t in ops.corr_bwd(x1, x2, go))
        o = ops.corr_fwd(x1, x2).clone()
        bad = badf = 0
        for i in range(RUNS):
            if i % 10 == 0:
                with torch.cuda.stream(side):
                    noise @ noise
            a, b = ops.corr_bwd(x1, x2, go)
            bad += 0 if (torch.equal(a, r1) and torch.equal(b, r2)) else 1
            badf += 0 if torch.equal(ops.corr_fwd(x1, x2), o) else 1
        total_bad += bad + badf
        print(f"corr {str(dtype)[6:]:9s} {shape}: bwd mismatching runs {bad}/{RUNS}, fwd {badf}/{RUNS}", flush=True)
for shape in [(8, 128, 8, 26), (8, 96, 16, 52)]:
    B, C, H, W = shape
    x, go = torch.randn(shape, generator=g).cuda(), torch.randn(shape, generator=g).cuda()
    fl = torch.randn(B, 2, H, W, generator=g).cuda()  # ~1 px: lists of <= 16 taps are sorted -> bit-identical
    rx, rf = (t.clone() for t in ops.warp_bwd(x, fl, go, "border"))
    bad = 0
    for i in range(RUNS):
        a, b = ops.warp_bwd(x, fl, go, "border")
        bad += 0 if (torch.equal(a, rx) and torch.equal(b, rf)) else 1
    total_bad += bad
    print(f"warp_bwd (fused small-image) {shape}: mismatching runs {bad}/{RUNS}", flush=True)
# segment max-pooling of the +mf branch: forward (integer atomicMax keys) and backward (fixed-order segment sums)
from unsamflow_b200 import segpool  # noqa: E402
for shape, n in [((8, 32, 64, 208), 40), ((8, 32, 16, 52), 256)]:
    B, C, H, W = shape
    feat, go = torch.randn(shape, generator=g).cuda(), torch.randn(shape, generator=g).cuda()
    ids = torch.randint(0, n, (B, H, W), generator=g, dtype=torch.int32).cuda()  # white-noise ids: worst case for the groups
    ro, ra = (t.clone() for t in segpool.segpool_fwd(feat, ids, n))
    rg = segpool.segpool_bwd(go, ids, ra, n).clone()
    bad = 0
    for i in range(RUNS):
        if i % 10 == 0:
            with torch.cuda.stream(side):
                noise @ noise
        o, a = segpool.segpool_fwd(feat, ids, n)
        bad += 0 if (torch.equal(o, ro) and torch.equal(a, ra) and torch.equal(segpool.segpool_bwd(go, ids, a, n), rg)) else 1
    total_bad += bad
    print(f"segpool {shape} {n} ids: mismatching runs {bad}/{RUNS}", flush=True)
torch.cuda.synchronize()
print("TOTAL mismatching runs:", total_bad)
sys.exit(1 if total_bad else 0)
