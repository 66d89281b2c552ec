#!/usr/bin/env python
"""Run selected hot-path ops a few times (for ncu / compute-sanitizer; prints nothing to judge).

  python benchmarks/profile_ops.py [--match corr_fwd.L4] [--reps 3] [--batch 8]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unsamflow_b200.hotpath import HotPathStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--match", default="")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--height", type=int, default=256)
    ap.add_argument("--width", type=int, default=832)
    args = ap.parse_args()
    step = HotPathStep(args.batch, args.height, args.width)
    ops = [op for op in step.ops if args.match in op.name and op.name.endswith("d12")]
    torch.cuda.synchronize()
    for _ in range(args.reps):
        for op in ops:
            op.fn()
    torch.cuda.synchronize()
    print("ran", [op.name for op in ops], "x", args.reps)


if __name__ == "__main__":
    main()
