#!/usr/bin/env python
"""Kernel timeline of the hot-path step as bench.py runs it (two direction streams + loss lanes, CUDA-graph replay,
L2 flushed before each step): start / duration / stream of every kernel from the CUPTI activity records that
torch.profiler collects.  Writes a compact JSON (one row per kernel of the last profiled step) and prints a summary:
per-kernel in-step durations, the busy fraction of each stream, and the critical chain.

    python benchmarks/step_timeline.py --out gpurun_out/timeline.json
"""
import argparse
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unsamflow_b200.hotpath import HotPathStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--height", type=int, default=256)
    ap.add_argument("--width", type=int, default=832)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default="")
    ap.add_argument("--batch-directions", action="store_true")
    args = ap.parse_args()
    step = HotPathStep(args.batch, args.height, args.width, batch_directions=args.batch_directions)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    run = (lambda: step.run()) if args.batch_directions else (lambda: step.run_concurrent(step.chains))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    for _ in range(10):
        flush.zero_()
        g.replay()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            flush.zero_()
            g.replay()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    rows = []
    for e in evs:
        rows.append({"name": e.name, "start_us": e.time_range.start, "dur_us": e.time_range.end - e.time_range.start,
                     "stream": getattr(e, "device_resource_id", None) if hasattr(e, "device_resource_id") else None})
    rows.sort(key=lambda r: r["start_us"])
    # split into steps at the flush kernels (vectorized fill of 256 MB: by far the longest memset-like kernel)
    fills = [i for i, r in enumerate(rows) if ("fill" in r["name"].lower() or "memset" in r["name"].lower()) and r["dur_us"] > 20]
    if len(fills) >= 2:
        last = rows[fills[-1] + 1:]
    else:
        last = rows
    t0 = min(r["start_us"] for r in last)
    for r in last:
        r["start_us"] = round(r["start_us"] - t0, 2)
        r["dur_us"] = round(r["dur_us"], 2)
        r["name"] = r["name"].replace("void usf::(anonymous namespace)::", "").replace("usf::(anonymous namespace)::", "")[:60]
    end = max(r["start_us"] + r["dur_us"] for r in last)
    print(f"last step: {len(last)} device activities, span {end:.1f} us")
    by = {}
    for r in last:
        k = r["name"].split("(")[0]
        by.setdefault(k, [0, 0.0])
        by[k][0] += 1
        by[k][1] += r["dur_us"]
    for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print(f"  {t:8.1f} us  x{n:3d}  {k}")
    streams = {}
    for r in last:
        streams.setdefault(r["stream"], 0.0)
        streams[r["stream"]] += r["dur_us"]
    print("busy us per stream:", {k: round(v, 1) for k, v in streams.items()})
    if args.out:
        with open(args.out, "w") as f:
            json.dump(last, f)


if __name__ == "__main__":
    main()
