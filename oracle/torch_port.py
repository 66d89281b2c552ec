"""PyTorch restatement of the reference's hot path (TEST INFRASTRUCTURE ONLY).

It drives the *same ATen operators in the same order* as the reference, so that
(a) its arithmetic is the reference's arithmetic on whatever device it runs on and
(b) timing it on the host cores is timing what the reference executes on a CPU
(``bench.py``: ``cpu_baseline.kind == "port"`` and ``--impl reference``).

Restates:
  cost_volume   models/correlation_native.py:13-23  (F.pad zeros, 81 shifted
                products, mean over channels, cat on dim 1)
  pixel_grid    utils/warp_utils.py:7-13   (mesh_grid: built on the CPU, integer)
  normalise     utils/warp_utils.py:16-23  (norm_grid: 2.0*v/(size-1) - 1.0, BHW2)
  flow_warp     utils/warp_utils.py:97-106 (grid_sample, align_corners=True)
  occ_mask_bidirection  utils/warp_utils.py:109-117
  range_map / occ_mask_backward  utils/warp_utils.py:26-94 / :120-126

Pinned against the reference's own outputs by tests/test_oracle_golden.py.
"""
import torch
import torch.nn.functional as F


def cost_volume(x1, x2, max_displacement=4):
    md = int(max_displacement)
    span = 2 * md + 1
    h, w = x1.shape[-2:]
    x2p = F.pad(x2, [md, md, md, md])
    planes = []
    for iy in range(span):
        for ix in range(span):
            window = x2p[:, :, iy:iy + h, ix:ix + w]
            planes.append(torch.mean(x1 * window, 1, keepdim=True))
    return torch.cat(planes, 1)


def pixel_grid(b, h, w):
    xs = torch.arange(0, w).repeat(b, h, 1)
    ys = torch.arange(0, h).repeat(b, w, 1).transpose(1, 2)
    return torch.stack([xs, ys], 1)


def normalise(v):
    h, w = v.shape[-2:]
    g = torch.zeros_like(v)
    g[:, 0] = 2.0 * v[:, 0] / (w - 1) - 1.0
    g[:, 1] = 2.0 * v[:, 1] / (h - 1) - 1.0
    return g.permute(0, 2, 3, 1)


def flow_warp(x, flow12, pad="border", mode="bilinear"):
    b, _, h, w = x.shape
    base = pixel_grid(b, h, w).type_as(x)
    grid = normalise(base + flow12)
    return F.grid_sample(x, grid, mode=mode, padding_mode=pad, align_corners=True)


def corr_fwd_bwd(x1, x2, gout, max_displacement=4):
    """Returns (out, gx1, gx2) through autograd, like the reference trains."""
    a = x1.detach().clone().requires_grad_(True)
    b = x2.detach().clone().requires_grad_(True)
    out = cost_volume(a, b, max_displacement)
    out.backward(gout)
    return out.detach(), a.grad, b.grad


def warp_fwd_bwd(x, flow, gout, pad="border", need_gx=True):
    a = x.detach().clone().requires_grad_(need_gx)
    f = flow.detach().clone().requires_grad_(True)
    out = flow_warp(a, f, pad=pad)
    out.backward(gout)
    return out.detach(), (a.grad if need_gx else None), f.grad


def occ_mask_bidirection(flow12, flow21, scale=0.01, bias=0.5):
    back = flow_warp(flow21, flow12, pad="zeros")
    diff = flow12 + back
    mag = (flow12 * flow12).sum(1, keepdim=True) + (back * back).sum(1, keepdim=True)
    return ((diff * diff).sum(1, keepdim=True) > scale * mag + bias).float()


def range_map(coords):
    """coords: [B,2,H,W] unnormalised target positions -> [B,1,H,W] sum of bilinear splat weights."""
    b, _, h, w = coords.shape
    x, y = coords[:, 0].reshape(b, -1), coords[:, 1].reshape(b, -1)
    xf, yf = torch.floor(x), torch.floor(y)
    xc, yc = xf + 1, yf + 1
    xf_in, yf_in = xf.clamp(0, w - 1), yf.clamp(0, h - 1)
    xc_in, yc_in = xc.clamp(0, w - 1), yc.clamp(0, h - 1)
    xc_out, yc_out, xf_out, yf_out = xc != xc_in, yc != yc_in, xf != xf_in, yf != yf_in
    taps = ((xc_in, yc_in, xc_out | yc_out), (xc_in, yf_in, xc_out | yf_out),
            (xf_in, yc_in, xf_out | yc_out), (xf_in, yf_in, xf_out | yf_out))
    index = torch.cat([tx + ty * w for tx, ty, _ in taps], 1).long()
    weight = torch.cat([(1 - torch.abs(x - tx)) * (1 - torch.abs(y - ty)) for tx, ty, _ in taps], 1)
    weight[torch.cat([bad for _, _, bad in taps], 1)] = 0
    acc = torch.zeros(b, h * w).type_as(coords)
    acc.scatter_add_(1, index, weight)
    return acc.view(b, h, w).unsqueeze(1)


def occ_mask_backward(flow21, th=0.2):
    b, _, h, w = flow21.shape
    base = pixel_grid(b, h, w).type_as(flow21)
    return (range_map(base + flow21).clamp(min=0.0, max=1.0) < th).float()
