"""GPU parity: the CUDA path (through the C ABI) against the reference's golden vectors and the
oracle.  Tolerance for fp32 (BASELINE.json north_star): 1e-5 relative, measured both as rel-L2 and
as max-abs error normalised by the reference's max-abs.

Nothing here reads /root/reference: fixtures are committed, the oracle is oracle/ (C + torch port).
"""
import numpy as np
import pytest
import torch

from conftest import max_norm_err, rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def usf():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import unsamflow_b200
    from unsamflow_b200 import ops
    unsamflow_b200.ops_mod = ops
    return unsamflow_b200


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def check(got, want, what, tol=TOL):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else got
    want = want.detach().cpu().numpy() if isinstance(want, torch.Tensor) else want
    assert got.shape == want.shape, what
    assert np.isfinite(got).all(), what
    e1, e2 = rel_l2(got, want), max_norm_err(got, want)
    assert e1 <= tol and e2 <= tol, f"{what}: rel_l2={e1:.3e} max_norm={e2:.3e} (tol {tol})"


# --------------------------------------------------------------------------------------------
# golden vectors produced by the unmodified reference
# --------------------------------------------------------------------------------------------
def test_corr_matches_reference_golden(usf, corr_golden):
    ops = usf.ops_mod
    for name, t in corr_golden.items():
        x1, x2, go = dev(t["x1"]), dev(t["x2"]), dev(t["gout"])
        check(ops.corr_fwd(x1, x2), t["out"], f"{name}.out")
        g1, g2 = ops.corr_bwd(x1, x2, go)
        check(g1, t["gx1"], f"{name}.gx1")
        check(g2, t["gx2"], f"{name}.gx2")


def test_warp_matches_reference_golden(usf, warp_golden):
    ops = usf.ops_mod
    for name, t in warp_golden.items():
        pad = str(t["pad"])
        x, fl, go = dev(t["x"]), dev(t["flow"]), dev(t["gout"])
        check(ops.warp_fwd(x, fl, pad), t["out"], f"{name}.out")
        for atomic in (False, True):
            gx, gfl = ops.warp_bwd(x, fl, go, pad, atomic=atomic)
            check(gx, t["gx"], f"{name}.gx(atomic={atomic})")
            check(gfl, t["gflow"], f"{name}.gflow(atomic={atomic})")


def test_autograd_modules_match_reference_golden(usf, corr_golden, warp_golden):
    t = corr_golden["c32_tile_edges"]
    x1, x2 = dev(t["x1"]).requires_grad_(True), dev(t["x2"]).requires_grad_(True)
    corr = usf.Correlation(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1,
                           corr_multiply=1)
    out = corr(x1, x2)
    check(out, t["out"], "module out")
    # the caller mutates the output in place (models/pwclite.py:308) before backward
    out2 = torch.nn.functional.leaky_relu(out, 0.1, inplace=True)
    go = dev(t["gout"])
    out2.backward(go)
    slope = torch.where(dev(t["out"]) > 0, 1.0, 0.1)
    from oracle import c_oracle
    g1, g2 = c_oracle.corr_bwd(t["x1"], t["x2"], (go * slope).cpu().numpy())
    check(x1.grad, g1, "module gx1 through in-place LeakyReLU")
    check(x2.grad, g2, "module gx2 through in-place LeakyReLU")

    t = warp_golden["border_c32"]
    x, fl = dev(t["x"]).requires_grad_(True), dev(t["flow"]).requires_grad_(True)
    y = usf.flow_warp(x, fl)  # positional call like models/pwclite.py:302
    y.backward(dev(t["gout"]))
    check(y, t["out"], "flow_warp out")
    check(x.grad, t["gx"], "flow_warp gx")
    check(fl.grad, t["gflow"], "flow_warp gflow")
    # loss-side call: keyword pad, gradient wrt flow only (losses/flow_loss.py:130)
    t = warp_golden["border_img_wide"]
    x, fl = dev(t["x"]), dev(t["flow"]).requires_grad_(True)
    usf.flow_warp(x, fl, pad="border").backward(dev(t["gout"]))
    check(fl.grad, t["gflow"], "flow_warp gflow only")


# --------------------------------------------------------------------------------------------
# oracle on seeded inputs: every pyramid level of the BASELINE configs, ragged / edge shapes
# --------------------------------------------------------------------------------------------
CORR_SHAPES = [
    # KITTI 256x832 pyramid, B=1 (configs[0]) and the B=8 bench sizes at the two ends
    (1, 192, 4, 13), (1, 128, 8, 26), (1, 96, 16, 52), (1, 64, 32, 104), (1, 32, 64, 208),
    (8, 192, 4, 13), (2, 32, 64, 208),
    # Sintel 384x832 and crops: odd widths, tile remainders
    (1, 128, 12, 26), (1, 64, 48, 104), (1, 192, 3, 10), (1, 96, 20, 44), (1, 192, 5, 11),
    # microbench sweep corner: C not a multiple of 8, C tiny, single pixel row/col
    (1, 196, 9, 33), (2, 3, 17, 70), (1, 1, 8, 32), (1, 40, 1, 5), (1, 33, 7, 1),
    # whole-image-in-shared-memory kernels (H*W <= 256): pair / single pixel variants, ragged channel groups,
    # more channels than one shared-memory chunk holds
    (3, 20, 6, 14), (2, 9, 16, 16), (8, 128, 8, 26), (2, 12, 4, 13), (1, 200, 8, 32), (2, 130, 15, 17), (17, 100, 8, 32),
]


@pytest.mark.parametrize("shape", CORR_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_corr_vs_oracle(usf, shape):
    from oracle import c_oracle
    ops = usf.ops_mod
    B, C, H, W = shape
    g = torch.Generator().manual_seed(1234 + C + H)
    x1, x2 = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
    go = torch.randn(B, 81, H, W, generator=g)
    out = ops.corr_fwd(x1.cuda(), x2.cuda())
    check(out, c_oracle.corr_fwd(x1.numpy(), x2.numpy()), "out")
    g1, g2 = ops.corr_bwd(x1.cuda(), x2.cuda(), go.cuda())
    o1, o2 = c_oracle.corr_bwd(x1.numpy(), x2.numpy(), go.numpy())
    check(g1, o1, "gx1")
    check(g2, o2, "gx2")


def test_corr_full_size_b8_vs_oracle(usf):
    """configs[1] largest level at the bench batch size, against the (multi-threaded) C oracle."""
    from oracle import c_oracle
    ops = usf.ops_mod
    shape = (8, 32, 64, 208)
    g = torch.Generator().manual_seed(7)
    x1, x2 = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
    go = torch.randn(8, 81, 64, 208, generator=g)
    check(ops.corr_fwd(x1.cuda(), x2.cuda()), c_oracle.corr_fwd(x1.numpy(), x2.numpy()), "out")
    g1, g2 = ops.corr_bwd(x1.cuda(), x2.cuda(), go.cuda())
    o1, o2 = c_oracle.corr_bwd(x1.numpy(), x2.numpy(), go.numpy())
    check(g1, o1, "gx1")
    check(g2, o2, "gx2")


def test_corr_nullable_grads_and_unaligned_views(usf):
    from oracle import c_oracle
    ops = usf.ops_mod
    g = torch.Generator().manual_seed(3)
    x1, x2 = torch.randn(2, 16, 12, 24, generator=g), torch.randn(2, 16, 12, 24, generator=g)
    go = torch.randn(2, 81, 12, 24, generator=g)
    o1, o2 = c_oracle.corr_bwd(x1.numpy(), x2.numpy(), go.numpy())
    g1, none = ops.corr_bwd(x1.cuda(), x2.cuda(), go.cuda(), need_gx2=False)
    assert none is None
    check(g1, o1, "gx1 only")
    none, g2 = ops.corr_bwd(x1.cuda(), x2.cuda(), go.cuda(), need_gx1=False)
    assert none is None
    check(g2, o2, "gx2 only")
    # the same on a whole-image (H*W <= 256) shape, where one CTA normally computes both gradients
    y1, y2 = torch.randn(2, 24, 8, 26, generator=g), torch.randn(2, 24, 8, 26, generator=g)
    gy = torch.randn(2, 81, 8, 26, generator=g)
    s1, s2 = c_oracle.corr_bwd(y1.numpy(), y2.numpy(), gy.numpy())
    check(ops.corr_bwd(y1.cuda(), y2.cuda(), gy.cuda(), need_gx2=False)[0], s1, "small gx1 only")
    check(ops.corr_bwd(y1.cuda(), y2.cuda(), gy.cuda(), need_gx1=False)[1], s2, "small gx2 only")
    # storage offset of 4 bytes -> the 16-byte vector path must not be taken
    buf = torch.zeros(x1.numel() + 1).cuda()
    buf[1:] = x1.flatten().cuda()
    x1u = buf[1:].view_as(x1)
    assert x1u.data_ptr() % 16 != 0
    check(ops.corr_fwd(x1u, x2.cuda()), c_oracle.corr_fwd(x1.numpy(), x2.numpy()), "unaligned out")
    # non-contiguous input (channels-last strides) is made contiguous at the boundary
    x2cl = x2.cuda().to(memory_format=torch.channels_last)
    check(ops.corr_fwd(x1.cuda(), x2cl), c_oracle.corr_fwd(x1.numpy(), x2.numpy()), "channels_last out")


def test_corr_fused_leaky_relu(usf):
    ops = usf.ops_mod
    g = torch.Generator().manual_seed(5)
    x1, x2 = torch.randn(1, 32, 16, 52, generator=g).cuda(), torch.randn(1, 32, 16, 52, generator=g).cuda()
    plain = ops.corr_fwd(x1, x2)
    fused = ops.corr_fwd(x1, x2, leaky_slope=0.1)
    assert torch.equal(fused, torch.nn.functional.leaky_relu(plain, 0.1))
    for shape in [(2, 128, 8, 26), (2, 192, 4, 13)]:
        y1, y2 = torch.randn(shape, generator=g).cuda(), torch.randn(shape, generator=g).cuda()
        assert torch.equal(ops.corr_fwd(y1, y2, leaky_slope=0.1), torch.nn.functional.leaky_relu(ops.corr_fwd(y1, y2), 0.1))


WARP_SHAPES = [
    # model-side warps (levels 1..4) and loss-side image warps of the KITTI / Sintel configs
    (1, 128, 8, 26), (1, 96, 16, 52), (1, 64, 32, 104), (2, 32, 64, 208), (1, 3, 256, 832),
    (1, 3, 48, 104), (1, 2, 64, 208), (1, 196, 5, 11), (1, 1, 2, 2), (2, 17, 9, 31),
]


@pytest.mark.parametrize("shape", [(2, 32, 64, 208), (2, 64, 32, 104), (2, 96, 16, 52), (2, 128, 8, 26), (2, 192, 4, 13),
                                   (1, 17, 9, 31), (2, 20, 40, 36), (3, 8, 70, 30)], ids=lambda s: "x".join(map(str, s)))
def test_corr_fwd_into_a_channel_slice(usf, shape):
    """usf_corr_fwd_into: the cost volume lands in channels [k, k+81) of a wider tensor (the estimator's concat buffer,
    models/pwclite.py:364-366), bit-identical to the dense result, nothing outside the slice touched -- every kernel
    family (TMA tiles, split-channel, whole-image, direct) and both epilogues."""
    ops = usf.ops_mod
    B, C, H, W = shape
    g = torch.Generator().manual_seed(40 + C)
    x1, x2 = torch.randn(shape, generator=g).cuda(), torch.randn(shape, generator=g).cuda()
    for slope in (None, 0.1):
        dense = ops.corr_fwd(x1, x2, leaky_slope=slope)
        for ctot, k in ((81 + 34, 0), (81 + 81 + 34, 81), (90, 5)):   # (90, 5): slice start not 16-byte aligned for odd H*W
            buf = torch.full((B, ctot, H, W), 7.0, device="cuda")
            ops.corr_fwd(x1, x2, leaky_slope=slope, out=buf[:, k:k + 81])
            assert torch.equal(buf[:, k:k + 81], dense), (slope, ctot, k)
            rest = torch.cat([buf[:, :k], buf[:, k + 81:]], 1)
            assert bool((rest == 7.0).all()), "wrote outside the slice"
    with pytest.raises(ValueError):
        ops.corr_fwd(x1, x2, out=torch.empty(B, 81, H, 2 * W, device="cuda")[..., ::2])
    hb = x1.bfloat16(), x2.bfloat16()
    dense = ops.corr_fwd(*hb, leaky_slope=0.1)
    buf = torch.zeros(B, 115, H, W, device="cuda", dtype=torch.bfloat16)
    ops.corr_fwd(*hb, leaky_slope=0.1, out=buf[:, :81])
    assert torch.equal(buf[:, :81], dense) and float(buf[:, 81:].abs().max()) == 0.0


def test_cost_concat_equals_corr_leaky_cat(usf):
    """correlation.cost_concat (kernel-side LeakyReLU + write into the concat buffer) against the reference's
    sequence corr -> LeakyReLU(inplace) -> torch.cat, values and gradients, one and two feature pairs."""
    from unsamflow_b200.correlation import cost_concat
    corr = usf.Correlation(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1)
    g = torch.Generator().manual_seed(50)
    B, C, H, W = 2, 32, 32, 104
    base = [torch.randn(B, C, H, W, generator=g).cuda() for _ in range(4)] + \
           [torch.randn(B, 32, H, W, generator=g).cuda(), torch.randn(B, 2, H, W, generator=g).cuda()]
    for n_pairs in (1, 2):
        a = [t.clone().requires_grad_(True) for t in base]
        b = [t.clone().requires_grad_(True) for t in base]
        pa = [(a[0], a[1]), (a[2], a[3])][:n_pairs]
        pb = [(b[0], b[1]), (b[2], b[3])][:n_pairs]
        fused = cost_concat(pa, [a[4], a[5]], 0.1)
        ref = torch.cat([torch.nn.functional.leaky_relu(corr(p, q), 0.1) for p, q in pb] + [b[4], b[5]], 1)
        assert torch.equal(fused, ref)
        go = torch.randn(ref.shape, generator=g).cuda()
        fused.backward(go)
        ref.backward(go)
        for i, (u, v) in enumerate(zip(a, b)):
            if i >= 2 * n_pairs and i < 4:
                assert u.grad is None and v.grad is None
                continue
            assert torch.allclose(u.grad, v.grad, rtol=1e-6, atol=1e-7), i


@pytest.mark.parametrize("pad", ["border", "zeros"])
@pytest.mark.parametrize("kind", ["gauss", "wild", "integer", "subpixel_edges"])
@pytest.mark.parametrize("shape", WARP_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_warp_vs_oracle(usf, shape, kind, pad):
    from oracle import c_oracle
    ops = usf.ops_mod
    B, C, H, W = shape
    g = torch.Generator().manual_seed(99 + C + H)
    x, go = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
    if kind == "gauss":
        fl = 3.0 * torch.randn(B, 2, H, W, generator=g)
    elif kind == "wild":
        fl = torch.randn(B, 2, H, W, generator=g) * torch.tensor([W / 4.0, H / 4.0]).view(1, 2, 1, 1)
    elif kind == "integer":
        fl = torch.randint(-3, 4, (B, 2, H, W), generator=g).float()
    else:  # coordinates landing exactly on / just around the image border
        fl = torch.zeros(B, 2, H, W)
        fl[:, 0] = -torch.arange(W).float().view(1, 1, W) + torch.tensor([0.0, -0.25, W - 1.0, W - 0.75])[
            torch.randint(0, 4, (B, H, W), generator=g)]
        fl[:, 1] = -torch.arange(H).float().view(1, H, 1) + torch.tensor([0.0, -0.25, H - 1.0, H - 0.75])[
            torch.randint(0, 4, (B, H, W), generator=g)]
    xn, fn, gn = x.numpy(), fl.numpy(), go.numpy()
    check(ops.warp_fwd(x.cuda(), fl.cuda(), pad), c_oracle.warp_fwd(xn, fn, pad), "out")
    ogx, ogf = c_oracle.warp_bwd(xn, fn, gn, pad)
    # "subpixel_edges" collapses the WHOLE image onto 16 target pixels: thousands of fp32 terms per
    # target, summed in hardware-scheduled order (the overflow side list / ATen-style atomics) against
    # a double-precision oracle -> allow fp32 summation noise there; everything else stays at 1e-5.
    gx_tol = 1e-4 if kind == "subpixel_edges" and H * W > 4096 else TOL
    for atomic in (False, True):
        gx, gf = ops.warp_bwd(x.cuda(), fl.cuda(), go.cuda(), pad, atomic=atomic)
        check(gx, ogx, f"gx(atomic={atomic})", tol=gx_tol)
        check(gf, ogf, f"gflow(atomic={atomic})")
    none, gf = ops.warp_bwd(x.cuda(), fl.cuda(), go.cuda(), pad, need_gx=False)
    assert none is None
    check(gf, ogf, "gflow only")
    gx, none = ops.warp_bwd(x.cuda(), fl.cuda(), go.cuda(), pad, need_gflow=False)
    assert none is None
    check(gx, ogx, "gx only", tol=gx_tol)


def test_warp_bwd_collapse_overflows_tap_lists(usf):
    """Every pixel flows far out of the image: border padding piles all taps onto the border
    pixels (far more than the 8 list slots each), exercising the overflow side list."""
    from oracle import c_oracle
    ops = usf.ops_mod
    # 16x40: the fused small-image kernel (16 shared-memory slots + shared side list); 40x40 and 2 images: the
    # three-kernel path (8 slots + global side list)
    for B, C, H, W in [(1, 8, 16, 40), (2, 8, 40, 40), (2, 40, 10, 24)]:
        g = torch.Generator().manual_seed(11)
        x, go = torch.randn(B, C, H, W, generator=g), torch.randn(B, C, H, W, generator=g)
        fl = torch.full((B, 2, H, W), 1000.0)
        fl[:, :, : H // 2] = torch.randn(B, 2, H // 2, W, generator=g) * 0.3 + 5.5  # pile onto few targets
        ogx, ogf = c_oracle.warp_bwd(x.numpy(), fl.numpy(), go.numpy(), "border")
        gx, gf = ops.warp_bwd(x.cuda(), fl.cuda(), go.cuda(), "border")
        check(gx, ogx, f"gx with overflow {H}x{W}")
        check(gf, ogf, f"gflow with overflow {H}x{W}")


def test_warp_bwd_persistent_workspace_stays_clean(usf):
    """USF_FLAG_WORKSPACE_CLEAN: a workspace zeroed once serves call after call without the per-call memset -- the
    gather zeroes the tap counters it has read and its last CTA resets the overflow counters -- also right after a
    call whose side list overflowed (collapse onto the border) and across different flows."""
    from oracle import c_oracle
    ops = usf.ops_mod
    B, C, H, W = 2, 8, 40, 44
    g = torch.Generator().manual_seed(21)
    x, go = torch.randn(B, C, H, W, generator=g), torch.randn(B, C, H, W, generator=g)
    collapse = torch.full((B, 2, H, W), 1000.0)
    collapse[:, :, : H // 2] = torch.randn(B, 2, H // 2, W, generator=g) * 0.3 + 5.5
    flows = [torch.randn(B, 2, H, W, generator=g), collapse, 4.0 * torch.randn(B, 2, H, W, generator=g), collapse,
             torch.zeros(B, 2, H, W)]
    ws = ops.warp_bwd_workspace(x.cuda())
    nclean = ops._lib.load().usf_warp_bwd_workspace_clean_bytes(B, C, H, W, 0)
    assert 0 < nclean < ws.numel()
    for i, fl in enumerate(flows):
        gx, gf = ops.warp_bwd(x.cuda(), fl.cuda(), go.cuda(), "border", workspace=ws, workspace_clean=True)
        ogx, ogf = c_oracle.warp_bwd(x.numpy(), fl.numpy(), go.numpy(), "border")
        check(gx, ogx, f"gx, call {i}")
        check(gf, ogf, f"gflow, call {i}")
        torch.cuda.synchronize()
        assert int(ws[:nclean].max()) == 0, f"counters not left clean after call {i}"
    with pytest.raises(ValueError):
        ops.warp_bwd(x.cuda(), flows[0].cuda(), go.cuda(), workspace_clean=True)


def test_warp_list_path_is_deterministic(usf):
    ops = usf.ops_mod
    g = torch.Generator().manual_seed(12)
    x = torch.randn(2, 32, 64, 208, generator=g).cuda()
    go = torch.randn(2, 32, 64, 208, generator=g).cuda()
    fl = (torch.randn(2, 2, 64, 208, generator=g) * 1.0).cuda()
    a, _ = ops.warp_bwd(x, fl, go, "zeros", need_gflow=False)
    b, _ = ops.warp_bwd(x, fl, go, "zeros", need_gflow=False)
    # lists are sorted by source pixel, so only targets with more than 8 taps (side list, atomics)
    # may differ between runs, and only in the last bits
    assert (a != b).float().mean().item() < 2e-2
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    # fused small-image kernel: lists of up to 16 taps are sorted -> bit-identical for a moderate flow
    x = torch.randn(4, 96, 16, 52, generator=g).cuda()
    go = torch.randn(4, 96, 16, 52, generator=g).cuda()
    fl = (torch.randn(4, 2, 16, 52, generator=g) * 1.0).cuda()
    a, fa = ops.warp_bwd(x, fl, go, "border")
    for _ in range(5):
        b, fb = ops.warp_bwd(x, fl, go, "border")
        assert torch.equal(a, b) and torch.equal(fa, fb)


# --------------------------------------------------------------------------------------------
# size-independent properties at the bench sizes
# --------------------------------------------------------------------------------------------
def test_corr_adjoint_identity_full_size(usf):
    """<gO, corr(x1,x2)> is bilinear: it equals <gx1, x1> and <gx2, x2> exactly in exact arithmetic."""
    ops = usf.ops_mod
    for shape in [(8, 32, 64, 208), (8, 64, 32, 104), (8, 96, 16, 52), (8, 128, 8, 26), (8, 192, 4, 13)]:
        B, C, H, W = shape
        g = torch.Generator().manual_seed(C)
        x1, x2 = torch.randn(shape, generator=g).cuda(), torch.randn(shape, generator=g).cuda()
        go = torch.randn(B, 81, H, W, generator=g).cuda()
        out = ops.corr_fwd(x1, x2)
        g1, g2 = ops.corr_bwd(x1, x2, go)
        lhs = (go.double() * out.double()).sum().item()
        scale = (go.double().abs() * out.double().abs()).sum().item()
        assert abs(lhs - (g1.double() * x1.double()).sum().item()) <= 1e-6 * scale
        assert abs(lhs - (g2.double() * x2.double()).sum().item()) <= 1e-6 * scale


def test_corr_swap_symmetry(usf):
    """corr(x1,x2)[k(d)](p) == corr(x2,x1)[k(-d)](p+d) wherever both are defined."""
    ops = usf.ops_mod
    g = torch.Generator().manual_seed(21)
    x1, x2 = torch.randn(2, 64, 32, 104, generator=g).cuda(), torch.randn(2, 64, 32, 104, generator=g).cuda()
    a, b = ops.corr_fwd(x1, x2), ops.corr_fwd(x2, x1)
    H, W = 32, 104
    for dy, dx in [(-4, -4), (0, 3), (2, -1), (4, 4)]:
        k, kneg = (dy + 4) * 9 + dx + 4, (-dy + 4) * 9 - dx + 4
        ys, xs = slice(max(0, -dy), min(H, H - dy)), slice(max(0, -dx), min(W, W - dx))
        yt = slice(ys.start + dy, ys.stop + dy)
        xt = slice(xs.start + dx, xs.stop + dx)
        assert torch.allclose(a[:, k, ys, xs], b[:, kneg, yt, xt], rtol=1e-5, atol=1e-6)


def test_warp_identity_and_integer_shift_full_size(usf):
    ops = usf.ops_mod
    g = torch.Generator().manual_seed(22)
    x = torch.randn(8, 3, 256, 832, generator=g).cuda()
    zero = torch.zeros(8, 2, 256, 832).cuda()
    y = ops.warp_fwd(x, zero, "border")
    assert max_norm_err(y.cpu().numpy(), x.cpu().numpy()) < 2e-4  # coordinate round trip: <= 6e-5 px at W=832
    fl = zero.clone()
    fl[:, 0] = 2.0
    fl[:, 1] = -1.0
    y = ops.warp_fwd(x, fl, "zeros")
    want = torch.zeros_like(x)
    want[:, :, 1:, :-2] = x[:, :, :-1, 2:]
    assert max_norm_err(y.cpu().numpy(), want.cpu().numpy()) < 2e-4


def test_warp_adjoint_identity_full_size(usf):
    """warp is linear in x: <gO, warp(x)> == <gx, x>."""
    ops = usf.ops_mod
    for shape, pad in [((8, 32, 64, 208), "border"), ((8, 3, 256, 832), "zeros")]:
        B, C, H, W = shape
        g = torch.Generator().manual_seed(C + H)
        x, go = torch.randn(shape, generator=g).cuda(), torch.randn(shape, generator=g).cuda()
        fl = (3.0 * torch.randn(B, 2, H, W, generator=g)).cuda()
        y = ops.warp_fwd(x, fl, pad)
        gx, _ = ops.warp_bwd(x, fl, go, pad, need_gflow=False)
        lhs = (go.double() * y.double()).sum().item()
        rhs = (gx.double() * x.double()).sum().item()
        scale = (go.double().abs() * y.double().abs()).sum().item()
        assert abs(lhs - rhs) <= 1e-6 * scale


# --------------------------------------------------------------------------------------------
# the torch port (what the reference runs) on the SAME device: ATen CUDA kernels as 2nd opinion
# --------------------------------------------------------------------------------------------
def test_against_torch_port_on_cuda(usf):
    """One decoder level (warp -> corr, fwd+bwd) against the ATen CUDA kernels the reference would
    run on this GPU.  ATen's CUDA `tensor / scalar` multiplies by the reciprocal while the CPU kernel
    divides, so both coordinate flavours are measured; the flavour matching the device must agree to
    the parity budget (the other one shows the reference's own CPU-vs-GPU spread)."""
    import json
    import os
    from oracle import torch_port
    g = torch.Generator().manual_seed(31)
    x1 = torch.randn(2, 64, 32, 104, generator=g).cuda().requires_grad_(True)
    x2 = torch.randn(2, 64, 32, 104, generator=g).cuda().requires_grad_(True)
    fl = (2.0 * torch.randn(2, 2, 32, 104, generator=g)).cuda().requires_grad_(True)
    go = torch.randn(2, 81, 32, 104, generator=g).cuda()
    corr = usf.Correlation(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1)
    ref = torch_port.cost_volume(x1, torch_port.flow_warp(x2, fl))
    ref.backward(go)
    want = [t.detach().cpu().numpy() for t in (ref, x1.grad, x2.grad, fl.grad)]
    errs = {}
    for recip in (False, True):
        for t in (x1, x2, fl):
            t.grad = None
        out = corr(x1, usf.flow_warp(x2, fl, coord_reciprocal=recip))
        out.backward(go)
        got = [t.detach().cpu().numpy() for t in (out, x1.grad, x2.grad, fl.grad)]
        errs["reciprocal" if recip else "divide"] = {
            n: max(rel_l2(a, b), max_norm_err(a, b)) for n, a, b in zip(["out", "gx1", "gx2", "gflow"], got, want)}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/coord_mode_vs_aten_cuda.json", "w") as f:
        json.dump(errs, f, indent=1)
    best = min(errs.values(), key=lambda e: max(e.values()))
    assert max(best.values()) <= 2e-5, errs


# --------------------------------------------------------------------------------------------
# boundary behaviour
# --------------------------------------------------------------------------------------------
def test_errors_are_loud(usf):
    ops = usf.ops_mod
    x = torch.zeros(1, 4, 8, 8).cuda()
    with pytest.raises(ValueError):
        ops.corr_fwd(x, torch.zeros(1, 4, 8, 9).cuda())
    with pytest.raises(RuntimeError):
        ops.corr_fwd(x, torch.zeros(1, 4, 8, 8))  # CPU tensor: no fallback
    with pytest.raises(ValueError):
        ops.warp_fwd(x, torch.zeros(1, 3, 8, 8).cuda())
    with pytest.raises(ValueError):
        ops.warp_fwd(torch.zeros(1, 4, 1, 8).cuda(), torch.zeros(1, 2, 1, 8).cuda())
    with pytest.raises(NotImplementedError):
        usf.flow_warp(x, torch.zeros(1, 2, 8, 8).cuda(), mode="nearest")
    with pytest.raises(NotImplementedError):
        ops.corr_fwd(x.half(), x.half())          # fp16 is not a supported storage type
    with pytest.raises(ValueError):
        ops.corr_fwd(x, x.bfloat16())
    # caller-supplied buffers reach the kernels as raw pointers: wrong shape / dtype / layout / device must raise
    fl, go = torch.zeros(1, 2, 8, 8).cuda(), torch.zeros(1, 81, 8, 8).cuda()
    with pytest.raises(ValueError, match="out=: shape"):
        ops.corr_fwd(x, x, out=torch.empty(1, 81, 8, 7).cuda())
    with pytest.raises(ValueError, match="out=: dtype"):
        ops.corr_fwd(x.bfloat16(), x.bfloat16(), out=torch.empty(1, 81, 8, 8).cuda())
    with pytest.raises(ValueError, match="contiguous"):
        ops.corr_fwd(x, x, out=torch.empty(1, 81, 8, 16).cuda()[..., ::2])
    with pytest.raises(ValueError, match="out=: device"):
        ops.corr_fwd(x, x, out=torch.empty(1, 81, 8, 8))
    with pytest.raises(ValueError, match="gx1=: shape"):
        ops.corr_bwd(x, x, go, gx1=torch.empty(1, 4, 8, 4).cuda())
    with pytest.raises(ValueError, match="gx2=: dtype"):
        ops.corr_bwd(x, x, go, gx2=torch.empty(1, 4, 8, 8, dtype=torch.bfloat16).cuda())
    with pytest.raises(ValueError, match="out=: shape"):
        ops.warp_fwd(x, fl, out=torch.empty(1, 4, 8, 9).cuda())
    with pytest.raises(ValueError, match="gx=: contiguous|gx=: must be contiguous"):
        ops.warp_bwd(x, fl, x, gx=torch.empty(1, 4, 8, 16).cuda()[..., ::2])
    with pytest.raises(ValueError, match="gflow=: dtype"):
        ops.warp_bwd(x, fl, x, gflow=torch.empty(1, 2, 8, 8, dtype=torch.bfloat16).cuda())
    big = torch.zeros(1, 4, 40, 40).cuda()
    with pytest.raises(ValueError, match="workspace"):
        ops.warp_bwd(big, torch.zeros(1, 2, 40, 40).cuda(), big, workspace=torch.empty(16, dtype=torch.uint8).cuda())
    with pytest.raises(ValueError, match="out=: shape"):
        ops.occ_bidirection(fl, fl, out=torch.empty(1, 2, 8, 8).cuda())
    with pytest.raises(ValueError, match="out=: dtype"):
        ops.occ_backward(fl, out=torch.empty(1, 1, 8, 8, dtype=torch.float64).cuda())


def test_cuda_graph_capture_of_a_step(usf):
    """The whole hot-path step is capturable (no allocation, no sync, no H2D inside)."""
    from unsamflow_b200.hotpath import HotPathStep
    step = HotPathStep(1, 64, 208, flow_sigma=1.5)
    step.run()
    torch.cuda.synchronize()
    eager = {k: v.clone() for k, v in step.t.items() if k.startswith(("corr.", "gx1.", "gflow.", "glflow."))}
    for v in eager:
        step.t[v].zero_()
    graph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step.run()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(graph):
        step.run()
    for v in eager:
        step.t[v].zero_()
    graph.replay()
    torch.cuda.synchronize()
    for k, v in eager.items():
        assert torch.allclose(step.t[k], v, rtol=1e-5, atol=1e-6), k


def test_fallback_kernels_in_a_subprocess(usf):
    """USF_DISABLE_TMA=1 + USF_DIRECT_MAX_PIXELS=0 + USF_SMALL_MAX_HW=0 force the cp.async kernels that normally
    only serve large odd-width shapes; USF_SMALL_MAX_HW=0 / USF_WARP_SMALL_MAX_HW=0 send the tiny levels to the direct corr kernels and the
    three-kernel warp backward; USF_BWD_MIXED_TILES=0 keeps gx1 on 8x32 tiles; USF_PDL=0 launches without programmatic
    stream serialization.
    Both must pass the same golden / oracle checks."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for extra in ({"USF_DISABLE_TMA": "1", "USF_DIRECT_MAX_PIXELS": "0", "USF_SMALL_MAX_HW": "0"},
                  {"USF_SMALL_MAX_HW": "0", "USF_WARP_SMALL_MAX_HW": "0"},
                  {"USF_BWD_MIXED_TILES": "0"}, {"USF_PDL": "0"}):
        env = dict(os.environ, **extra)
        sel = "golden or corr_vs_oracle or nullable"
        if "USF_WARP_SMALL_MAX_HW" in extra:  # small images through the three-kernel warp backward
            sel += " or warp_vs_oracle or collapse"
        if "USF_BWD_MIXED_TILES" in extra:    # both gradients on 8x32 tiles at the full-size level
            sel = "corr_full_size or adjoint"
        if "USF_PDL" in extra:                # plain launches instead of programmatic dependent launches
            sel = "golden or graph_capture"
        r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu",
                            os.path.join(here, "test_gpu_parity.py"), "-k", sel],
                           env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, str(extra) + r.stdout[-3000:] + r.stderr[-2000:]


def test_corr_kernels_are_bit_deterministic_under_load(usf):
    """50 back-to-back runs at the bench size must be bit-identical (guards the TMA / mbarrier pipelines:
    a stage released before its loads had landed once showed up as rare, large errors in gx2)."""
    ops = usf.ops_mod
    g = torch.Generator().manual_seed(7)
    # mixed-tile TMA backward, 8x32 tiles, the 4-channel thin-grid form at 3 CTAs/SM (where the stage-release
    # race resurfaced before release_stage_reads() got its fence: 5 of 300 runs differed), whole-image kernels
    for shape, runs in [((8, 32, 64, 208), 50), ((8, 64, 32, 104), 50), ((8, 96, 16, 52), 150),
                        ((1, 64, 32, 104), 150), ((8, 128, 8, 26), 50)]:
        B, C, H, W = shape
        x1, x2 = torch.randn(shape, generator=g).cuda(), torch.randn(shape, generator=g).cuda()
        go = torch.randn(B, 81, H, W, generator=g).cuda()
        o0 = ops.corr_fwd(x1, x2).clone()
        a0, b0 = (t.clone() for t in ops.corr_bwd(x1, x2, go))
        for _ in range(runs):
            a, b = ops.corr_bwd(x1, x2, go)
            assert torch.equal(a, a0) and torch.equal(b, b0), shape
            assert torch.equal(ops.corr_fwd(x1, x2), o0), shape


# --------------------------------------------------------------------------------------------
# bf16 storage (new capability): fp32 accumulate, fp32 flow / coordinates.
# Stated tolerance: rel-L2 <= 4e-3 and max-norm <= 8e-3 against the fp32 oracle evaluated on the
# SAME bf16-rounded inputs (i.e. one bf16 rounding of each output, 2^-9 relative).
# --------------------------------------------------------------------------------------------
BF16_L2, BF16_MAX = 4e-3, 8e-3


def check_bf16(got, want, what):
    got = got.detach().float().cpu().numpy()
    assert np.isfinite(got).all(), what
    e1, e2 = rel_l2(got, want), max_norm_err(got, want)
    assert e1 <= BF16_L2 and e2 <= BF16_MAX, f"{what}: rel_l2={e1:.3e} max_norm={e2:.3e}"


@pytest.mark.parametrize("shape", [(2, 64, 8, 26), (1, 33, 5, 11), (2, 32, 32, 104), (1, 96, 16, 52),
                                   (8, 32, 64, 208), (2, 64, 48, 104), (1, 196, 24, 40), (3, 20, 9, 72),
                                   (8, 196, 24, 40), (8, 20, 9, 72), (2, 40, 30, 88)],
                         ids=lambda s: "x".join(map(str, s)))
def test_bf16_corr_vs_oracle(usf, shape):
    from oracle import c_oracle
    ops = usf.ops_mod
    B, C, H, W = shape
    g = torch.Generator().manual_seed(C + W)
    x1 = torch.randn(shape, generator=g).bfloat16()
    x2 = torch.randn(shape, generator=g).bfloat16()
    go = torch.randn(B, 81, H, W, generator=g).bfloat16()
    out = ops.corr_fwd(x1.cuda(), x2.cuda())
    assert out.dtype == torch.bfloat16
    check_bf16(out, c_oracle.corr_fwd(x1.float().numpy(), x2.float().numpy()), "out")
    g1, g2 = ops.corr_bwd(x1.cuda(), x2.cuda(), go.cuda())
    o1, o2 = c_oracle.corr_bwd(x1.float().numpy(), x2.float().numpy(), go.float().numpy())
    check_bf16(g1, o1, "gx1")
    check_bf16(g2, o2, "gx2")
    only1, none = ops.corr_bwd(x1.cuda(), x2.cuda(), go.cuda(), need_gx2=False)
    assert none is None and torch.equal(only1, g1)
    none, only2 = ops.corr_bwd(x1.cuda(), x2.cuda(), go.cuda(), need_gx1=False)
    assert none is None and torch.equal(only2, g2)


@pytest.mark.parametrize("pad", ["border", "zeros"])
@pytest.mark.parametrize("shape", [(2, 32, 16, 52), (1, 3, 64, 208), (1, 17, 9, 31)], ids=lambda s: "x".join(map(str, s)))
def test_bf16_warp_vs_oracle(usf, shape, pad):
    from oracle import c_oracle
    ops = usf.ops_mod
    B, C, H, W = shape
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(shape, generator=g).bfloat16()
    go = torch.randn(shape, generator=g).bfloat16()
    fl = 2.0 * torch.randn(B, 2, H, W, generator=g)
    xn, gn, fn = x.float().numpy(), go.float().numpy(), fl.numpy()
    out = ops.warp_fwd(x.cuda(), fl.cuda(), pad)
    assert out.dtype == torch.bfloat16
    check_bf16(out, c_oracle.warp_fwd(xn, fn, pad), "out")
    gx, gf = ops.warp_bwd(x.cuda(), fl.cuda(), go.cuda(), pad)
    ogx, ogf = c_oracle.warp_bwd(xn, fn, gn, pad)
    assert gx.dtype == torch.bfloat16 and gf.dtype == torch.float32
    check_bf16(gx, ogx, "gx")
    check(gf, ogf, "gflow (fp32 accumulate over bf16 storage)")


def test_bf16_autograd_decoder_level(usf):
    """warp -> corr in bf16 through the public modules (autocast-style caller)."""
    from oracle import c_oracle
    g = torch.Generator().manual_seed(5)
    x1 = torch.randn(1, 32, 16, 52, generator=g).bfloat16()
    x2 = torch.randn(1, 32, 16, 52, generator=g).bfloat16()
    fl = 1.5 * torch.randn(1, 2, 16, 52, generator=g)
    a, b, f = x1.cuda().requires_grad_(True), x2.cuda().requires_grad_(True), fl.cuda().requires_grad_(True)
    out = usf.Correlation()(a, usf.flow_warp(b, f))
    out.float().sum().backward()
    x2w = c_oracle.warp_fwd(x2.float().numpy(), fl.numpy(), "border")
    x2w_b = torch.from_numpy(x2w).bfloat16().float().numpy()      # the warp output is stored in bf16
    check_bf16(out, c_oracle.corr_fwd(x1.float().numpy(), x2w_b), "out")
    assert a.grad.dtype == torch.bfloat16 and f.grad.dtype == torch.float32
    assert torch.isfinite(a.grad.float()).all() and torch.isfinite(f.grad).all()


def test_empty_batch_and_nonfinite_flow(usf):
    ops = usf.ops_mod
    e = torch.zeros(0, 8, 6, 12).cuda()
    assert tuple(ops.corr_fwd(e, e).shape) == (0, 81, 6, 12)
    assert tuple(usf.flow_warp(e, torch.zeros(0, 2, 6, 12).cuda()).shape) == (0, 8, 6, 12)
    # NaN / Inf / astronomically large flows must not fault and must only poison their own pixels
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 4, 8, 16, generator=g).cuda()
    go = torch.randn(1, 4, 8, 16, generator=g).cuda()
    fl = torch.zeros(1, 2, 8, 16).cuda()
    fl[0, 0, 0, 0], fl[0, 1, 1, 1], fl[0, 0, 2, 2], fl[0, 1, 3, 3] = float("nan"), float("inf"), 1e30, -1e30
    for pad in ("zeros", "border"):
        y = ops.warp_fwd(x, fl, pad)
        gx, gf = ops.warp_bwd(x, fl, go, pad)
        clean = torch.ones(8, 16, dtype=torch.bool).cuda()
        for (r, c) in ((0, 0), (1, 1), (2, 2), (3, 3)):
            clean[r, c] = False
        assert torch.isfinite(y[0, :, clean]).all() and torch.isfinite(gf[0, :, clean]).all()
        assert torch.allclose(y[0, :, clean], x[0, :, clean], atol=1e-5)   # zero flow elsewhere: identity
