#!/usr/bin/env python
"""torch.profiler summary of the PWC-Lite training step (where the non-hot-path time goes)."""
import os, sys
import torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unsamflow_b200.train_step import TrainStep
torch.backends.cudnn.benchmark = True
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ts = TrainStep(B, 256, 832, "cuda:0", batch_directions=True)
for _ in range(3):
    ts.step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5):
    ts.step()
torch.cuda.synchronize()
print("B", B, "ms/step wall", (time.perf_counter() - t0) / 5 * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        ts.step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
