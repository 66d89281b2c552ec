#!/usr/bin/env python
"""BASELINE.json configs[3] (kitti_aug+hg+mf): forward + backward of PWCLite(add_mask_corr=True) on synthetic 256x832
pairs with SAM-shaped segment maps (piecewise-constant ids), B pairs, both directions batched.  Three poolings of
the mask features (models/pwclite.py:317-340): the usf kernels (csrc/segpool.cu), the PyTorch scatter formulation
(pwclite.segment_max_pool_torch) and the reference's one-hot product (may run out of memory: reported as null).
Device time per iteration from CUDA events, median of 5.  Prints one JSON line."""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unsamflow_b200.pwclite as pw  # noqa: E402
from unsamflow_b200 import segment_max_pool  # noqa: E402


def onehot_pool(feat, seg, n=None):
    oh = torch.nn.functional.one_hot(seg[:, 0].long())[:, None]
    pooled = torch.amax(oh * feat[..., None], dim=(2, 3))
    return (oh * pooled[:, :, None, None, :]).sum(dim=-1)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    ids = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    torch.backends.cudnn.benchmark = True
    gen = torch.Generator().manual_seed(0)
    img1, img2 = torch.rand(B, 3, 256, 832, generator=gen).cuda(), torch.rand(B, 3, 256, 832, generator=gen).cuda()
    coarse = torch.randint(0, ids, (B, 1, 16, 52), generator=gen).float()
    s1 = torch.nn.functional.interpolate(coarse, (256, 832), mode="nearest").cuda()
    s2 = torch.roll(s1, 9, dims=-1)
    out = {"config": {"batch": B, "ids": ids, "height": 256, "width": 832}}
    for name, pool in (("usf_segpool", lambda f, s, n=None: segment_max_pool(f, s, 256)),
                       ("torch_scatter", lambda f, s, n=None: pw.segment_max_pool_torch(f, s)), ("one_hot", onehot_pool)):
        pw.segment_max_pool = pool
        torch.manual_seed(1)
        model = pw.PWCLite(batch_directions=True, add_mask_corr=True, fuse_cost_concat=True).cuda()
        model.init_weights()

        def it():
            for p in model.parameters():
                p.grad = None
            o = model(img1, img2, s1, s2, with_bk=True)
            sum(f.abs().mean() for f in o["flows_12"] + o["flows_21"]).backward()
        try:
            torch.cuda.reset_peak_memory_stats()
            for _ in range(2):
                it()
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); it(); b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b))
            out[name] = {"ms_per_iter": round(statistics.median(ts), 2),
                         "peak_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2)}
        except torch.cuda.OutOfMemoryError:
            out[name] = None
        del model
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
