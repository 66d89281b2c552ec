"""GPU parity of the mask-feature pooling kernels (csrc/segpool.cu) against the reference's own
formulation -- one_hot x feature -> amax -> one_hot-weighted sum, models/pwclite.py:319-328, evaluated by PyTorch in
this test -- and against its autograd.  Values must be bit-identical (a maximum involves no rounding); gradients are
segment sums in another order: 1e-5 relative."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def usf():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import unsamflow_b200
    return unsamflow_b200


def onehot_pool(feat, seg, n=None):
    """models/pwclite.py:319-328 verbatim in spirit: F.one_hot(seg.long()), amax of the product, weighted sum."""
    oh = torch.nn.functional.one_hot(seg[:, 0].long(), -1 if n is None else n)[:, None]   # [B,1,h,w,N]
    pooled = torch.amax(oh * feat[..., None], dim=(2, 3))
    return (oh * pooled[:, :, None, None, :]).sum(dim=-1)


def sam_like_ids(B, h, w, n, gen, cell=4):
    coarse = torch.randint(0, n, (B, 1, (h + cell - 1) // cell, (w + cell - 1) // cell), generator=gen).float()
    return torch.nn.functional.interpolate(coarse, (h, w), mode="nearest")


CASES = [
    # (B, C, h, w, ids, kind)
    (2, 32, 16, 52, 40, "sam"), (2, 32, 64, 208, 40, "sam"), (1, 32, 8, 26, 40, "sam"), (2, 5, 6, 7, 9, "random"),
    (1, 3, 9, 31, 200, "random"), (2, 32, 32, 104, 256, "sam"), (1, 7, 4, 13, 1, "one"), (2, 6, 10, 12, 30, "gaps"),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c[:5])) + c[5])
@pytest.mark.parametrize("sign", ["mixed", "negative"])
def test_pooling_matches_the_one_hot_formulation(usf, case, sign):
    B, C, h, w, n, kind = case
    gen = torch.Generator().manual_seed(7 + C + h)
    feat = torch.randn(B, C, h, w, generator=gen)
    if sign == "negative":
        feat = -feat.abs() - 0.1  # every in-segment value is negative: the reference pools to 0 (zeros from outside)
    if kind == "sam":
        seg = sam_like_ids(B, h, w, n, gen)
    elif kind == "one":
        seg = torch.zeros(B, 1, h, w)  # one segment covers the image: no zero competes, the true (negative) maximum
    else:
        seg = torch.randint(0, n, (B, 1, h, w), generator=gen).float()
        if kind == "gaps":
            seg = seg * 3
    a = feat.cuda().requires_grad_(True)
    b = feat.cuda().requires_grad_(True)
    mine = usf.segment_max_pool(a, seg.cuda())
    ref = onehot_pool(b, seg.cuda())
    assert torch.equal(mine, ref)
    g = torch.randn(mine.shape, generator=gen).cuda()
    mine.backward(g)
    ref.backward(g)
    scale = b.grad.abs().max().clamp_min(1e-30)
    assert ((a.grad - b.grad).abs().max() / scale).item() <= 1e-5
    # a fixed class count (no device->host sync) gives the same result
    again = usf.segment_max_pool(feat.cuda(), seg.cuda(), num_segments=int(seg.max()) + 5)
    assert torch.equal(again, ref.detach())


def test_pooling_is_bit_reproducible_and_capturable(usf):
    gen = torch.Generator().manual_seed(1)
    feat = torch.randn(8, 32, 64, 208, generator=gen).cuda().requires_grad_(True)
    seg = sam_like_ids(8, 64, 208, 40, gen).cuda()
    g = torch.randn(8, 32, 64, 208, generator=gen).cuda()
    outs, grads = [], []
    for _ in range(4):
        feat.grad = None
        o = usf.segment_max_pool(feat, seg, num_segments=256)
        o.backward(g)
        outs.append(o.detach().clone())
        grads.append(feat.grad.clone())
    assert all(torch.equal(outs[0], o) for o in outs) and all(torch.equal(grads[0], x) for x in grads)
    # mass conservation of the gradient: every segment's gradient sum lands on exactly one pixel (or on the clamp)
    from unsamflow_b200 import segpool
    ids = seg[:, 0].to(torch.int32).contiguous()
    out, arg = segpool.segpool_fwd(feat.detach(), ids, 256)
    gf = segpool.segpool_bwd(g, ids, arg, 256)
    won = arg >= 0                                            # [B,C,N]
    seg_sum = torch.zeros(8, 32, 256, device="cuda").scatter_add_(2, ids.view(8, 1, -1).expand(8, 32, -1).long(),
                                                                   g.view(8, 32, -1))
    assert torch.allclose(gf.sum(dim=(2, 3)), (seg_sum * won).sum(-1), rtol=1e-4, atol=1e-3)
    # sync-free with a fixed class count: capturable in a CUDA graph
    static_out = torch.empty_like(out)
    graph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        segpool.segpool_fwd(feat.detach(), ids, 256, out=static_out)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(graph):
        segpool.segpool_fwd(feat.detach(), ids, 256, out=static_out)
    static_out.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_out, outs[0])


def test_pooling_bf16_and_errors(usf):
    gen = torch.Generator().manual_seed(2)
    feat = torch.randn(2, 32, 16, 52, generator=gen).cuda().bfloat16()
    seg = sam_like_ids(2, 16, 52, 40, gen).cuda()
    mine = usf.segment_max_pool(feat, seg)
    assert mine.dtype == torch.bfloat16 and torch.equal(mine.float(), onehot_pool(feat.float(), seg))
    with pytest.raises(RuntimeError, match="no CPU path"):
        usf.segment_max_pool(feat.cpu(), seg.cpu())
    with pytest.raises(ValueError):
        usf.segment_max_pool(feat, seg[:, :, :8])
    with pytest.raises(ValueError, match="num_segments"):
        usf.segment_max_pool(feat, seg, num_segments=100000)
    # ids outside [0, num_segments) belong to no segment: zeros out, no gradient
    f32 = feat.float().requires_grad_(True)
    o = usf.segment_max_pool(f32, seg + 1000.0, num_segments=8)
    assert float(o.abs().max()) == 0.0
    o.sum().backward()
    assert float(f32.grad.abs().max()) == 0.0


def test_pooling_speed_against_the_one_hot_formulation(usf):
    """Recorded, not asserted tightly: the kernel pair against the reference's formulation at the largest level."""
    import time
    gen = torch.Generator().manual_seed(3)
    feat = torch.randn(8, 32, 64, 208, generator=gen).cuda().requires_grad_(True)
    seg = sam_like_ids(8, 64, 208, 40, gen).cuda()
    g = torch.randn(8, 32, 64, 208, generator=gen).cuda()

    def timed(fn, reps=10):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e6

    def ours():
        feat.grad = None
        usf.segment_max_pool(feat, seg, num_segments=40).backward(g)

    def ref():
        feat.grad = None
        onehot_pool(feat, seg, 40).backward(g)

    t_ours, t_ref = timed(ours), timed(ref)
    print(f"segment pooling fwd+bwd 8x32x64x208, 40 ids: usf {t_ours:.0f} us, one-hot formulation {t_ref:.0f} us")
    assert t_ours < t_ref
