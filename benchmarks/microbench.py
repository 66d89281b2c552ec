#!/usr/bin/env python
"""Per-op microbenchmark (BASELINE.json configs[4]): CUDA-event timing of single ops through the C ABI.

  python benchmarks/microbench.py --ops corr_fwd,corr_bwd --shapes 8x32x64x208,8x64x32x104 [--cold]

Each (op, shape) is captured in a CUDA graph holding `--inner` back-to-back launches so that Python
launch overhead does not pollute the number; --cold flushes L2 before every graph replay (then
inner is forced to 1).  Prints one JSON line per (op, shape) and an optional table file.
"""
import argparse
import json
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unsamflow_b200 import hotpath, ops  # noqa: E402

FLOW_KIND = "smooth"
DTYPE = torch.float32
KITTI = ["8x192x4x13", "8x128x8x26", "8x96x16x52", "8x64x32x104", "8x32x64x208"]


def build(op, shape, sigma):
    B, C, H, W = shape
    g = torch.Generator().manual_seed(1)
    dev = "cuda"
    x1 = torch.randn(B, C, H, W, generator=g).to(dev).to(DTYPE)
    x2 = torch.randn(B, C, H, W, generator=g).to(dev).to(DTYPE)
    if op == "corr_fwd":
        out = torch.empty(B, 81, H, W, device=dev, dtype=DTYPE)
        return (lambda: ops.corr_fwd(x1, x2, out=out)), hotpath.corr_fwd_bytes(*shape), hotpath.corr_fwd_flops(*shape), 1
    if op == "corr_bwd":
        go = torch.randn(B, 81, H, W, generator=g).to(dev).to(DTYPE)
        g1, g2 = torch.empty_like(x1), torch.empty_like(x2)
        return (lambda: ops.corr_bwd(x1, x2, go, gx1=g1, gx2=g2)), hotpath.corr_bwd_bytes(*shape), 2 * hotpath.corr_fwd_flops(*shape), 2
    fl = (hotpath.synthetic_flow(B, H, W, sigma, g) if FLOW_KIND == "smooth"
          else sigma * torch.randn(B, 2, H, W, generator=g)).to(dev)
    if op == "warp_fwd":
        out = torch.empty_like(x1)
        return (lambda: ops.warp_fwd(x1, fl, "border", out=out)), hotpath.warp_fwd_bytes(*shape), 0, 1
    if op in ("warp_bwd", "warp_bwd_atomic"):
        atomic = op.endswith("atomic")
        gx, gf = torch.empty_like(x1), torch.empty_like(fl)
        ws = None if atomic else torch.empty(ops.warp_bwd_workspace_bytes(x1), dtype=torch.uint8, device=dev)
        return (lambda: ops.warp_bwd(x1, fl, x2, "border", gx=gx, gflow=gf, workspace=ws, atomic=atomic)), \
            hotpath.warp_bwd_bytes(*shape), 0, (1 if atomic else 3)
    if op == "warp_bwd_flow":
        gf = torch.empty_like(fl)
        return (lambda: ops.warp_bwd(x1, fl, x2, "border", need_gx=False, gflow=gf)), \
            hotpath.warp_bwd_bytes(*shape, need_gx=False), 0, 1
    raise ValueError(op)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="corr_fwd,corr_bwd,warp_fwd,warp_bwd,warp_bwd_flow")
    ap.add_argument("--shapes", default=",".join(KITTI))
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--inner", type=int, default=10)
    ap.add_argument("--cold", action="store_true")
    ap.add_argument("--sigma", type=float, default=3.0)
    ap.add_argument("--out", default="")
    ap.add_argument("--flow", default="smooth", choices=["smooth", "iid"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"], help="storage type of feature tensors")
    ap.add_argument("--preset", default="", choices=["", "sweep"],
                    help="sweep: BASELINE.json configs[4] (C=32..196, every pyramid level, B=1..64)")
    args = ap.parse_args()
    global FLOW_KIND, DTYPE
    FLOW_KIND = args.flow
    DTYPE = {"f32": torch.float32, "bf16": torch.bfloat16}[args.dtype]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    # wake the GPU up: a fresh box idles at ~120 MHz and takes a while to reach its boost clock
    burn = torch.randn(4096, 4096, device="cuda")
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 1.0:
        burn @ burn
    torch.cuda.synchronize()
    try:
        import pynvml
        pynvml.nvmlInit()
        nvh = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
    except Exception:  # noqa: BLE001
        pynvml, nvh = None, None
    inner = 1 if args.cold else args.inner
    work = [(op, s) for op in args.ops.split(",") for s in args.shapes.split(",")]
    if args.preset == "sweep":
        lv = [(192, 4, 13), (128, 8, 26), (96, 16, 52), (64, 32, 104), (32, 64, 208)]
        shapes = [f"{b}x{c}x{h}x{w}" for (c, h, w) in lv for b in (1, 8, 64)]
        shapes += [f"8x{c}x64x208" for c in (64, 96, 128, 192, 196)]                 # channel sweep at the largest level
        shapes += ["4x32x96x208", "4x64x48x104", "4x32x112x256", "8x32x48x160", "8x32x80x176"]  # Sintel / crops
        work = [(op, s) for op in ("corr_fwd", "corr_bwd", "warp_fwd", "warp_bwd") for s in shapes]
        work += [(op, s) for op in ("warp_fwd", "warp_bwd_flow") for s in ("8x3x256x832", "4x3x384x832", "4x3x448x1024")]
    rows = []
    for op, s in work:
        if True:
            shape = tuple(int(v) for v in s.split("x"))
            fn, nbytes, flops, launches = build(op, shape, args.sigma)
            for _ in range(3):
                fn()
            for _ in range(20):
                burn @ burn
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for _ in range(inner):
                    fn()
            ts = []
            for _ in range(args.iters):
                if args.cold:
                    flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); graph.replay(); b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b) * 1e3 / inner)
            us = statistics.median(ts)
            mhz = pynvml.nvmlDeviceGetClockInfo(nvh, pynvml.NVML_CLOCK_SM) if nvh is not None else None
            row = {"op": op, "shape": s, "us": round(us, 2), "min_us": round(min(ts), 2), "launches": launches,
                   "GBps": round(nbytes / us / 1e3, 1), "TFLOPs": round(flops / us / 1e6, 2),
                   "l2": "cold" if args.cold else "warm", "sm_mhz": mhz, "dtype": args.dtype}
            rows.append(row)
            print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
