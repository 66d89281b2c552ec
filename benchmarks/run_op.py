#!/usr/bin/env python
"""Run one op at one shape a few times (ncu target).  python benchmarks/run_op.py corr_fwd 16x32x64x208 3"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from microbench import build
op, shape, reps = sys.argv[1], tuple(int(v) for v in sys.argv[2].split("x")), int(sys.argv[3])
fn, *_ = build(op, shape, 3.0)
for _ in range(reps):
    fn()
torch.cuda.synchronize()
print("ok")
