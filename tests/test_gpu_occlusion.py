"""GPU parity of the two occlusion-mask kernels (through the C ABI) against the reference's golden vectors
and the oracle (utils/warp_utils.py:109-126).

The masks are comparisons of fp32 sums.  Parity bar: bit-exact, except that a pixel may differ where the
oracle's own margin (|lhs - threshold|) is at fp32 rounding level -- such a pixel flips between two correct fp32
evaluations of the reference too.  The range map underneath the backward mask is held to 2e-6 relative
(fp32 summation noise of the reference's scatter_add_; the kernel's sum is exact).

Nothing here reads /root/reference: fixtures are committed, the oracle is oracle/ (C + torch port).
"""
import numpy as np
import pytest
import torch

from conftest import max_norm_err, rel_l2

pytestmark = pytest.mark.gpu

MARGIN = 1e-5   # relative to max(1, |threshold side|)
RANGE_TOL = 2e-6


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from unsamflow_b200 import ops as o
    return o


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def real_mismatches(got, want, margin, size=1.0):
    return int(((got != want) & (np.abs(margin) > MARGIN * np.maximum(size, 1.0))).sum())


def smooth_flow(B, H, W, amp, seed):
    gen = torch.Generator().manual_seed(seed)
    coarse = amp * torch.randn(B, 2, max(H // 16, 2), max(W // 16, 2), generator=gen)
    return torch.nn.functional.interpolate(coarse, (H, W), mode="bilinear", align_corners=True).contiguous()


def check_backward(ops, flow21, want_occ, want_range, what):
    occ, rng = ops.occ_backward(dev(flow21), want_range_map=True)
    occ, rng = host(occ), host(rng)
    assert occ.shape == want_occ.shape == rng.shape, what
    assert set(np.unique(occ)) <= {0.0, 1.0}, what
    assert rel_l2(rng, want_range) <= RANGE_TOL and max_norm_err(rng, want_range) <= RANGE_TOL, \
        (what, rel_l2(rng, want_range), max_norm_err(rng, want_range))
    assert real_mismatches(occ, want_occ, want_range - np.float32(0.2)) == 0, what
    # the mask alone (no range map requested) is the same mask
    assert np.array_equal(host(ops.occ_backward(dev(flow21))), occ), what
    return int((occ != want_occ).sum())


def check_bidirection(ops, flow12, flow21, want_occ, margin, what, **kw):
    occ = host(ops.occ_bidirection(dev(flow12), dev(flow21), **kw))
    assert occ.shape == want_occ.shape, what
    assert set(np.unique(occ)) <= {0.0, 1.0}, what
    mag = (flow12.astype(np.float64) ** 2).sum(1, keepdims=True)
    assert real_mismatches(occ, want_occ, margin, 0.5 + 0.02 * mag) == 0, what
    return int((occ != want_occ).sum())


# ---- golden vectors produced by the unmodified reference ---------------------------------------------
def test_masks_match_reference_golden(ops, occ_golden):
    from oracle import c_oracle
    flips = 0
    for name, t in occ_golden.items():
        flips += check_backward(ops, t["flow21"], t["occ_backward"], t["range_map"], f"{name}.backward")
        _, margin = c_oracle.occ_bidirection(t["flow12"], t["flow21"])
        flips += check_bidirection(ops, t["flow12"], t["flow21"], t["occ_bidirection"], margin,
                                   f"{name}.bidirection")
    assert flips <= 3, flips  # threshold ties are rare: essentially bit-exact on the fixtures


# ---- oracle at the sizes the loss uses (kitti_base: finest flows 256x832, B = 8) --------------------------
@pytest.mark.parametrize("shape,amp", [((8, 256, 832), 6.0), ((2, 384, 832), 3.0), ((3, 37, 61), 2.0),
                                       ((1, 2, 2), 0.5), ((2, 5, 3), 40.0)])
def test_masks_match_oracle(ops, shape, amp):
    from oracle import c_oracle
    B, H, W = shape
    flow12 = smooth_flow(B, H, W, amp, 11)
    gen = torch.Generator().manual_seed(12)
    flow21 = (-flow12 + 0.5 * torch.randn(B, 2, H, W, generator=gen)).contiguous()
    f12, f21 = flow12.numpy(), flow21.numpy()
    want_occ, want_rng = c_oracle.occ_backward(f21)
    flips = check_backward(ops, f21, want_occ, want_rng, f"backward {shape}")
    want_occ, margin = c_oracle.occ_bidirection(f12, f21)
    assert 0.02 < want_occ.mean() < 0.98 or H * W < 100  # a mixed mask, or the test shows nothing
    flips += check_bidirection(ops, f12, f21, want_occ, margin, f"bidirection {shape}")
    assert flips <= max(3, B * H * W // 200000), flips
    # non-default thresholds (the reference's keyword arguments)
    want_occ, want_rng = c_oracle.occ_backward(f21, th=0.6)
    assert real_mismatches(host(ops.occ_backward(dev(f21), th=0.6)), want_occ, want_rng - np.float32(0.6)) == 0
    want_occ, margin = c_oracle.occ_bidirection(f12, f21, scale=0.05, bias=1.5)
    check_bidirection(ops, f12, f21, want_occ, margin, "bidirection kwargs", scale=0.05, bias=1.5)


def test_against_torch_port_on_cuda(ops):
    """The reference's own op sequence executed by ATen on this GPU (what the reference computes when it trains)."""
    from oracle import torch_port
    B, H, W = 4, 128, 416
    torch.manual_seed(22)
    flow12 = smooth_flow(B, H, W, 5.0, 21).cuda()
    flow21 = (-flow12 + 0.5 * torch.randn(B, 2, H, W, device="cuda")).contiguous()
    want = torch_port.occ_mask_backward(flow21)
    base = torch_port.pixel_grid(B, H, W).type_as(flow21)
    rng = torch_port.range_map(base + flow21)
    got, got_rng = ops.occ_backward(flow21, want_range_map=True)
    assert rel_l2(host(got_rng), host(rng)) <= RANGE_TOL
    assert real_mismatches(host(got), host(want), host(rng) - np.float32(0.2)) == 0
    want = host(torch_port.occ_mask_bidirection(flow12, flow21))
    got = host(ops.occ_bidirection(flow12, flow21, reciprocal=True))  # ATen-CUDA divides by reciprocal multiply
    # no margin is available from the port, and ATen's bilinear sum differs from ours by ~5e-7 relative: only
    # pixels within that distance of the threshold may differ (a handful in 213k)
    assert (got != want).mean() < 1e-4


# ---- size-independent properties at full size -------------------------------------------------------------
def test_range_map_properties_full_size(ops):
    B, H, W = 8, 256, 832
    zero = torch.zeros(B, 2, H, W, device="cuda")
    occ, rng = ops.occ_backward(zero, want_range_map=True)
    assert torch.equal(rng, torch.ones_like(rng)) and not occ.any()      # every pixel receives exactly its own 1
    occ, rng = ops.occ_backward(zero + 5000.0, want_range_map=True)
    assert occ.all() and not rng.any()                                    # everything leaves the image
    # integer translation: the range map is a shifted indicator, exactly
    shift = zero.clone()
    shift[:, 0] = 3.0
    shift[:, 1] = -2.0
    occ, rng = ops.occ_backward(shift, want_range_map=True)
    want = torch.zeros_like(rng)
    want[:, :, : H - 2, 3:] = 1.0
    assert torch.equal(rng, want) and torch.equal(occ, 1 - want)
    # mass conservation: sub-pixel motion that keeps every tap inside -> the map sums to the pixel count
    flow = smooth_flow(B, H, W, 0.3, 31).clamp(-0.45, 0.45).cuda()
    flow[:, :, 0, :] = 0
    flow[:, :, -1, :] = 0
    flow[:, :, :, 0] = 0
    flow[:, :, :, -1] = 0
    _, rng = ops.occ_backward(flow, want_range_map=True)
    assert abs(rng.double().sum().item() - B * H * W) < 1e-3 * B
    # bidirection: consistent flows (flow21 = -flow12, constant) are visible in the interior, and a mask that
    # disagrees everywhere is fully occluded
    const = zero.clone()
    const[:, 0] = 4.0
    occ = ops.occ_bidirection(const, -const)
    assert not occ[:, :, :, : W - 4].any()          # the last 4 columns warp from outside (zeros pad) and fail
    assert occ[:, :, :, W - 3:].all()
    assert ops.occ_bidirection(const, const).all()


def test_collapse_onto_one_pixel_is_exact(ops):
    """Every source pixel lands on (2.25, 3.5): 425,984 taps on four targets through the integer atomics."""
    B, H, W = 2, 256, 832
    ys, xs = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    flow = torch.stack([2.25 - xs, 3.5 - ys]).expand(B, 2, H, W).contiguous().cuda()
    occ, rng = ops.occ_backward(flow, want_range_map=True)
    n = H * W
    want = torch.zeros_like(rng)
    want[:, 0, 3, 2] = n * 0.75 * 0.5
    want[:, 0, 3, 3] = n * 0.25 * 0.5
    want[:, 0, 4, 2] = n * 0.75 * 0.5
    want[:, 0, 4, 3] = n * 0.25 * 0.5
    assert torch.equal(rng, want)               # the fixed-point sum is exact; a float scatter is not
    assert occ.sum().item() == B * (n - 4)


def test_non_finite_flow_contributes_nothing(ops):
    flow = torch.zeros(1, 2, 16, 20, device="cuda")
    flow[0, 0, 3, 3] = float("nan")
    flow[0, 1, 4, 4] = float("inf")
    flow[0, 0, 5, 5] = -float("inf")
    flow[0, 1, 6, 6] = 3e38
    occ, rng = ops.occ_backward(flow, want_range_map=True)
    assert torch.isfinite(rng).all()
    assert rng.sum().item() == 16 * 20 - 4 and occ.sum().item() == 4
    got = ops.occ_bidirection(flow, flow)   # NaN compares false -> "visible", like the reference's `>`
    assert set(np.unique(host(got))) <= {0.0, 1.0}


def test_bit_determinism_and_graph_capture(ops):
    B, H, W = 8, 256, 832
    torch.manual_seed(42)
    flow12 = smooth_flow(B, H, W, 6.0, 41).cuda()
    flow21 = (-flow12 + 0.5 * torch.randn(B, 2, H, W, device="cuda")).contiguous()
    first = ops.occ_backward(flow21, want_range_map=True)
    for _ in range(20):
        again = ops.occ_backward(flow21, want_range_map=True)
        assert torch.equal(again[0], first[0]) and torch.equal(again[1], first[1])
    bi = ops.occ_bidirection(flow12, flow21)
    # capture: caller-owned outputs and workspace, no allocation / sync / H2D inside
    out_b, out_bi = torch.empty_like(first[0]), torch.empty_like(bi)
    ws = torch.empty(ops.occ_backward_workspace_bytes(flow21), dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ops.occ_backward(flow21, out=out_b, workspace=ws)
        ops.occ_bidirection(flow12, flow21, out=out_bi)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ops.occ_backward(flow21, out=out_b, workspace=ws)
        ops.occ_bidirection(flow12, flow21, out=out_bi)
    for _ in range(3):
        out_b.fill_(7.0)
        out_bi.fill_(7.0)
        ws.fill_(255)      # the workspace needs no initialisation by the caller
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out_b, first[0]) and torch.equal(out_bi, bi)


def test_drop_in_signatures_and_errors(ops):
    import unsamflow_b200 as usf
    from oracle import torch_port
    torch.manual_seed(52)
    flow12 = smooth_flow(2, 64, 208, 3.0, 51).cuda()
    flow21 = (-flow12 + 0.4 * torch.randn(2, 2, 64, 208, device="cuda")).requires_grad_(True)
    m1 = usf.get_occu_mask_backward(flow21, th=0.2)                # utils/warp_utils.py:120
    m2 = usf.get_occu_mask_bidirection(flow12, flow21, 0.01, 0.5)  # utils/warp_utils.py:109
    assert m1.shape == m2.shape == (2, 1, 64, 208) and m1.dtype == torch.float32
    assert not m1.requires_grad and not m2.requires_grad
    assert (m1 != torch_port.occ_mask_backward(flow21.detach())).float().mean() < 1e-4  # threshold ties only
    # slices of the [B,4,H,W] flow the loss passes (losses/flow_loss.py:101-102) are non-contiguous views
    both = torch.cat([flow12, flow21.detach()], 1)
    assert torch.equal(usf.get_occu_mask_backward(both[:, 2:]), m1)
    with pytest.raises(RuntimeError):
        usf.get_occu_mask_backward(flow12.cpu())
    with pytest.raises(ValueError):
        ops.occ_bidirection(flow12, flow21[:1])
    with pytest.raises(ValueError):
        ops.occ_backward(flow12[:, :1])
    # the photometric loss runs on the native masks and backpropagates through the warps
    from unsamflow_b200.flow_loss import UnFlowLoss
    img1, img2 = torch.rand(2, 3, 64, 208, device="cuda"), torch.rand(2, 3, 64, 208, device="cuda")
    flows = [torch.cat([flow12, flow21], 1)] + [
        torch.nn.functional.interpolate(both, scale_factor=0.5 ** i, mode="bilinear") for i in (1, 2, 3, 4)]
    loss = UnFlowLoss()(flows, img1, img2)
    loss.backward()
    assert torch.isfinite(loss) and torch.isfinite(flow21.grad).all()


SPLAT_VARIANTS = {
    "unmerged": ({"USF_OCC_MERGE": "0", "USF_OCC_ITER": "1"}, True),
    "iter4": ({"USF_OCC_ITER": "4"}, True),
    "iter4_unmerged": ({"USF_OCC_MERGE": "0", "USF_OCC_ITER": "4"}, False),
    "tile_window": ({"USF_OCC_TILE": "1"}, True),
}


@pytest.mark.parametrize("variant", sorted(SPLAT_VARIANTS))
def test_splat_variants_in_subprocesses_are_bit_identical(ops, tmp_path, variant):
    """USF_OCC_MERGE=0 issues all four atomics per pixel instead of folding x-neighbouring lanes first;
    USF_OCC_ITER selects 1 or 4 pixels per thread; USF_OCC_TILE=1 accumulates per source tile in a shared-memory
    window first.  Every add is an integer add of the same fixed-point terms: each variant must give the same
    bits as the default, and pass the same golden / property checks."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    wild = smooth_flow(1, 70, 150, 4.0, 63)
    wild[:, :, 20:40, 30:90] *= 9.0   # a region whose taps leave the tile window
    flows = {"smooth": smooth_flow(4, 128, 416, 5.0, 61), "wild_patch": wild.contiguous(),
             "rough": 3.0 * torch.randn(2, 2, 61, 37, generator=torch.Generator().manual_seed(62))}
    torch.save(flows, tmp_path / "flows.pt")
    mine = {k: ops.occ_backward(v.cuda(), want_range_map=True)[1].cpu() for k, v in flows.items()}
    script = ("import sys, torch; sys.path.insert(0, %r); from unsamflow_b200 import ops; "
              "f = torch.load(%r); "
              "torch.save({k: ops.occ_backward(v.cuda(), want_range_map=True)[1].cpu() for k, v in f.items()}, %r)"
              % (os.path.dirname(here), str(tmp_path / "flows.pt"), str(tmp_path / "other.pt")))
    extra, rerun = SPLAT_VARIANTS[variant]
    env = dict(os.environ, **extra)
    if rerun:
        r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.abspath(__file__),
                            "-k", "golden or collapse or non_finite or properties"],
                           env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, str(extra) + r.stdout[-3000:] + r.stderr[-2000:]
    r = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, str(extra) + r.stdout[-2000:] + r.stderr[-2000:]
    other = torch.load(tmp_path / "other.pt")
    for k in flows:
        assert torch.equal(mine[k], other[k]), (extra, k)
