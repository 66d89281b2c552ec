#!/usr/bin/env python
"""Segment max-pooling of the +mf branch (csrc/segpool.cu) against the reference's one-hot formulation
(models/pwclite.py:319-328) on the same GPU, at the decoder levels of a KITTI step with add_mask_corr (C = 32 after the
1x1 projection, B = 8, both images = 16 maps), for 40 (typical SAM map) and 256 (uint8 id range) classes.
Device time from CUDA events, median of 20, warm L2; the one-hot formulation includes its [B,C,h,w,N] temporaries
(peak bytes reported).  Prints one JSON document."""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unsamflow_b200 import segpool  # noqa: E402


def onehot_pool(feat, seg, n):
    oh = torch.nn.functional.one_hot(seg.long(), n)[:, None]
    pooled = torch.amax(oh * feat[..., None], dim=(2, 3))
    return (oh * pooled[:, :, None, None, :]).sum(dim=-1)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return round(statistics.median(ts), 1)


def main():
    g = torch.Generator().manual_seed(0)
    rows = []
    for (h, w) in [(4, 13), (8, 26), (16, 52), (32, 104), (64, 208)]:
        for n in (40, 256):
            B, C = 16, 32
            feat = torch.randn(B, C, h, w, generator=g).cuda()
            go = torch.randn(B, C, h, w, generator=g).cuda()
            coarse = torch.randint(0, n, (B, 1, max(1, h // 4), max(1, w // 4)), generator=g).float()
            ids = torch.nn.functional.interpolate(coarse, (h, w), mode="nearest")[:, 0].to(torch.int32).cuda().contiguous()
            out, arg = segpool.segpool_fwd(feat, ids, n)
            row = {"shape": [B, C, h, w], "ids": n,
                   "usf_fwd_us": timed(lambda: segpool.segpool_fwd(feat, ids, n, out=out)),
                   "usf_bwd_us": timed(lambda: segpool.segpool_bwd(go, ids, arg, n))}
            try:
                fr = feat.clone().requires_grad_(True)
                torch.cuda.reset_peak_memory_stats()
                base = torch.cuda.memory_allocated()

                def ref():
                    fr.grad = None
                    onehot_pool(fr, ids, n).backward(go)
                row["onehot_fwd_bwd_us"] = timed(ref, reps=5)
                row["onehot_peak_MB"] = round((torch.cuda.max_memory_allocated() - base) / 1e6, 1)
                assert torch.equal(onehot_pool(feat, ids, n), out)
            except torch.cuda.OutOfMemoryError:
                row["onehot_fwd_bwd_us"] = None
            rows.append(row)
            print(json.dumps(row), flush=True)
    print(json.dumps({"segpool_compare": rows}))


if __name__ == "__main__":
    main()
