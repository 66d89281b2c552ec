"""The reference's OWN CUDA kernels (models/correlation_package, compiled for sm_100 from where they lie
by oracle/build_ref_cuda.py into the git-ignored oracle/_ref/) against ours on the same B200:
parity, and a timing table written to gpurun_out/ref_cuda_compare.json (tests are the only place that
may execute oracle/_ref besides bench.py's baseline legs)."""
import importlib.util
import json
import os
import statistics

import pytest
import torch

from conftest import ROOT, max_norm_err, rel_l2

pytestmark = pytest.mark.gpu
SO = os.path.join(ROOT, "oracle", "_ref", "correlation_cuda.so")


@pytest.fixture(scope="module")
def ref_cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/correlation_cuda.so not built (python oracle/build_ref_cuda.py in the build container)")
    spec = importlib.util.spec_from_file_location("correlation_cuda", SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_fwd(mod, x1, x2):
    rb1, rb2, out = x1.new_empty(0), x1.new_empty(0), x1.new_empty(0)
    mod.forward(x1, x2, rb1, rb2, out, 4, 1, 4, 1, 1, 1)     # correlation.py:24
    return out


def ref_bwd(mod, x1, x2, go):
    rb1, rb2, g1, g2 = (x1.new_empty(0) for _ in range(4))
    mod.backward(x1, x2, rb1, rb2, go, g1, g2, 4, 1, 4, 1, 1, 1)  # correlation.py:56
    return g1, g2


def _t(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)


def test_parity_and_speed_vs_reference_cuda_package(ref_cuda):
    from unsamflow_b200 import ops
    rows = []
    for shape in [(8, 192, 4, 13), (8, 128, 8, 26), (8, 96, 16, 52), (8, 64, 32, 104), (8, 32, 64, 208)]:
        B, C, H, W = shape
        g = torch.Generator().manual_seed(C)
        x1, x2 = torch.randn(shape, generator=g).cuda(), torch.randn(shape, generator=g).cuda()
        go = torch.randn(B, 81, H, W, generator=g).cuda()
        r_out = ref_fwd(ref_cuda, x1, x2)
        r_g1, r_g2 = ref_bwd(ref_cuda, x1, x2, go)
        out = ops.corr_fwd(x1, x2)
        g1, g2 = ops.corr_bwd(x1, x2, go)
        for a, b, n in ((out, r_out, "out"), (g1, r_g1, "gx1"), (g2, r_g2, "gx2")):
            a, b = a.cpu().numpy(), b.cpu().numpy()
            assert rel_l2(a, b) <= 1e-5 and max_norm_err(a, b) <= 1e-5, (shape, n, rel_l2(a, b))
        o, e1, e2 = torch.empty_like(out), torch.empty_like(x1), torch.empty_like(x2)
        rows.append({"shape": list(shape),
                     "ref_fwd_us": round(_t(lambda: ref_fwd(ref_cuda, x1, x2)), 1),
                     "usf_fwd_us": round(_t(lambda: ops.corr_fwd(x1, x2, out=o)), 1),
                     "ref_bwd_us": round(_t(lambda: ref_bwd(ref_cuda, x1, x2, go)), 1),
                     "usf_bwd_us": round(_t(lambda: ops.corr_bwd(x1, x2, go, gx1=e1, gx2=e2)), 1)})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ref_cuda_compare.json"), "w") as f:
        json.dump({"note": "eager calls, CUDA events, median of 20, includes Python launch overhead on both sides",
                   "rows": rows}, f, indent=1)
