"""Host-side mirror of the reference interface: constructor / call-signature behaviour, loud
failures without a GPU, and the hot-path shape and byte accounting."""
import inspect

import pytest
import torch

import unsamflow_b200
from unsamflow_b200 import hotpath


def test_correlation_ctor_matches_reference_call_site():
    # models/pwclite.py:208-215
    m = unsamflow_b200.Correlation(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1,
                                   corr_multiply=1)
    assert (m.pad_size, m.kernel_size, m.max_displacement, m.stride1, m.stride2) == (4, 1, 4, 1, 1)
    assert m.output_dim == 9 and not list(m.parameters())
    # correlation_native.py:7 style (extra kwargs swallowed)
    unsamflow_b200.Correlation(max_displacement=4, kernel_size=1, stride1=1, stride2=1, corr_multiply=1)
    for bad in (dict(max_displacement=3), dict(kernel_size=3), dict(stride2=2), dict(pad_size=3)):
        with pytest.raises(NotImplementedError):
            unsamflow_b200.Correlation(**bad)


def test_flow_warp_signature_matches_reference():
    # utils/warp_utils.py:97: flow_warp(x, flow12, pad="border", mode="bilinear")
    params = list(inspect.signature(unsamflow_b200.flow_warp).parameters.values())
    assert [p.name for p in params[:4]] == ["x", "flow12", "pad", "mode"]
    assert params[2].default == "border" and params[3].default == "bilinear"
    with pytest.raises(NotImplementedError):
        unsamflow_b200.flow_warp(torch.zeros(1, 1, 2, 2), torch.zeros(1, 2, 2, 2), mode="nearest")
    with pytest.raises(ValueError):
        unsamflow_b200.flow_warp(torch.zeros(1, 1, 2, 2), torch.zeros(1, 2, 2, 2), pad="reflection")


def test_no_cpu_fallback():
    x = torch.zeros(1, 2, 4, 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        unsamflow_b200.Correlation()(x, x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        unsamflow_b200.flow_warp(x, torch.zeros(1, 2, 4, 4))


def test_occlusion_masks_match_reference_signatures_and_fail_loudly_on_cpu():
    # utils/warp_utils.py:109: get_occu_mask_bidirection(flow12, flow21, scale=0.01, bias=0.5)
    params = list(inspect.signature(unsamflow_b200.get_occu_mask_bidirection).parameters.values())
    assert [(p.name, p.default) for p in params] == [("flow12", inspect.Parameter.empty), ("flow21", inspect.Parameter.empty),
                                                     ("scale", 0.01), ("bias", 0.5)]
    # utils/warp_utils.py:120: get_occu_mask_backward(flow21, th=0.2)
    params = list(inspect.signature(unsamflow_b200.get_occu_mask_backward).parameters.values())
    assert [(p.name, p.default) for p in params] == [("flow21", inspect.Parameter.empty), ("th", 0.2)]
    flow = torch.zeros(1, 2, 4, 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        unsamflow_b200.get_occu_mask_backward(flow)
    with pytest.raises(RuntimeError, match="no CPU path"):
        unsamflow_b200.get_occu_mask_bidirection(flow, flow)
    with pytest.raises(ValueError):
        unsamflow_b200.get_occu_mask_backward(torch.zeros(1, 3, 4, 4))
    # the loss takes its visibility masks from the kernel: on a CPU tensor it must raise, not fall back
    from unsamflow_b200.flow_loss import UnFlowLoss
    with pytest.raises(RuntimeError, match="no CPU path"):
        UnFlowLoss()([torch.zeros(1, 4, 8, 8)], torch.zeros(1, 3, 8, 8), torch.zeros(1, 3, 8, 8))


def test_preallocated_buffers_are_validated():
    """ops._check_out / _check_ws guard every caller-supplied buffer (they reach the kernels as raw pointers)."""
    from unsamflow_b200 import ops
    dev = torch.device("cpu")
    good = torch.empty(2, 3, 4, 5)
    assert ops._check_out(good, (2, 3, 4, 5), torch.float32, dev, "out") is good
    assert ops._check_out(None, (1,), torch.float32, dev, "out") is None
    for bad, pat in ((torch.empty(2, 3, 4, 6), "shape"), (torch.empty(2, 3, 4, 5, dtype=torch.bfloat16), "dtype"),
                     (torch.empty(2, 3, 4, 10)[..., ::2], "contiguous"), ("not a tensor", "expected a tensor")):
        with pytest.raises(ValueError, match=pat):
            ops._check_out(bad, (2, 3, 4, 5), torch.float32, dev, "out")
    with pytest.raises(ValueError, match="device"):
        ops._check_out(good, (2, 3, 4, 5), torch.float32, torch.device("cuda", 0), "out")
    ws = torch.empty(64, dtype=torch.uint8)
    assert ops._check_ws(ws, 64, dev) is ws
    with pytest.raises(ValueError, match="need 65"):
        ops._check_ws(ws, 65, dev)
    with pytest.raises(ValueError, match="aligned"):
        ops._check_ws(torch.empty(128, dtype=torch.uint8)[1:], 64, dev)


def test_round2_callers_fail_loudly_on_cpu_and_validate_arguments():
    """segment_max_pool (models/pwclite.py:317-340) and cost_concat (:307-308, 364-366) have no CPU path either."""
    from unsamflow_b200 import segment_max_pool
    from unsamflow_b200.correlation import cost_concat
    feat, seg = torch.zeros(1, 4, 6, 8), torch.zeros(1, 1, 6, 8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        segment_max_pool(feat, seg)
    with pytest.raises(ValueError):
        segment_max_pool(torch.zeros(4, 6, 8), seg)
    with pytest.raises(RuntimeError, match="no CPU path"):
        cost_concat([(feat, feat)], [torch.zeros(1, 2, 6, 8)])
    # signature of the reference-facing entry: segment_max_pool(feat, seg, num_segments=None)
    params = list(inspect.signature(segment_max_pool).parameters.values())
    assert [(p.name, p.default) for p in params] == [("feat", inspect.Parameter.empty), ("seg", inspect.Parameter.empty),
                                                     ("num_segments", None)]
    # out= of corr_fwd may be a channel slice of a wider contiguous tensor, nothing else that is non-contiguous
    from unsamflow_b200 import ops
    buf = torch.empty(2, 115, 6, 8)
    dev = buf.device
    assert ops._is_channel_slice(buf[:, 3:84], (2, 81, 6, 8), torch.float32, dev)
    assert not ops._is_channel_slice(buf[:, 3:84, :, ::2], (2, 81, 6, 4), torch.float32, dev)
    assert not ops._is_channel_slice(buf[:, 3:84].transpose(2, 3), (2, 81, 8, 6), torch.float32, dev)
    assert not ops._is_channel_slice(buf[:, 3:84], (2, 81, 6, 8), torch.bfloat16, dev)


def test_product_never_imports_the_oracle():
    import os
    import re
    root = os.path.dirname(unsamflow_b200.__file__)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "usf_oracle" not in text, f


def test_pyramid_shapes_match_survey():
    assert hotpath.pyramid_shapes(256, 832) == [(192, 4, 13), (128, 8, 26), (96, 16, 52), (64, 32, 104),
                                                (32, 64, 208)]
    assert hotpath.pyramid_shapes(384, 832) == [(192, 6, 13), (128, 12, 26), (96, 24, 52), (64, 48, 104),
                                                (32, 96, 208)]
    assert hotpath.pyramid_shapes(448, 1024)[0] == (192, 7, 16)
    assert hotpath.pyramid_shapes(192, 640)[0] == (192, 3, 10)
    assert hotpath.pyramid_shapes(320, 704)[-1] == (32, 80, 176)
    assert hotpath.loss_shapes(256, 832) == [(3, 256, 832), (3, 128, 416), (3, 64, 208), (3, 32, 104)]


def test_algorithmic_bytes_match_baseline_md():
    # KITTI pyramid B=1 11.8 / 17.8 MB, B=8 94.3 / 142.7 MB; Sintel B=4 70.7 / 107.0 MB
    def tot(fn, B, H, W):
        return sum(fn(B, *s) for s in hotpath.pyramid_shapes(H, W)) / 1e6
    assert round(tot(hotpath.corr_fwd_bytes, 1, 256, 832), 1) == 11.8
    assert round(tot(hotpath.corr_bwd_bytes, 1, 256, 832), 1) == 17.8
    assert round(tot(hotpath.corr_fwd_bytes, 8, 256, 832), 1) == 94.3
    assert round(tot(hotpath.corr_bwd_bytes, 8, 256, 832), 1) == 142.7
    assert round(tot(hotpath.corr_fwd_bytes, 4, 384, 832), 1) == 70.7
    assert round(tot(hotpath.corr_bwd_bytes, 4, 384, 832), 1) == 107.0
    assert hotpath.warp_fwd_bytes(1, 32, 64, 208) == (2 * 32 + 2) * 4 * 64 * 208
    assert hotpath.warp_bwd_bytes(1, 3, 256, 832, need_gx=False) == (2 * 3 + 4) * 4 * 256 * 832
    assert round(tot(hotpath.corr_fwd_flops, 8, 256, 832) / 1e3, 2) == 0.98  # GFLOP
