"""PWC-Lite with the B200-native hot path wired in: the CALLER of Correlation / flow_warp.

Scope note: the product of this repo is the correlation + warp path; this
module restates the network around it (models/pwclite.py:14-434, base configuration:
reduce_dense=True, learned_upsampler=True; optionally the mask-feature branch of the `+mf` configs,
add_mask_corr=True, configs/kitti_aug+hg+mf.json:3-6) only so that the path can be exercised and timed inside a
real training step without the reference tree on the box.  All learned layers are plain PyTorch / cuDNN,
exactly as in the reference; the call sites are

    x2_warp = flow_warp(x2, flow)                                 models/pwclite.py:302
    out_corr = self.corr(x1, x2_warp)                             models/pwclite.py:307
    x_mask_feat2_warp = flow_warp(x_mask_feat2, flow)             models/pwclite.py:355   (add_mask_corr)
    out_mask_corr = self.corr(x_mask_feat1, x_mask_feat2_warp)    models/pwclite.py:356   (add_mask_corr)

The segment pooling in front of the second pair (models/pwclite.py:317-340) runs as a segmented max + gather in
hand-written kernels (unsamflow_b200/segpool.py, csrc/segpool.cu) instead of the reference's one-hot product, whose
[B,32,h,w,N] temporaries reach gigabytes at B=8; same values and gradients
(`segment_max_pool_torch` is the PyTorch formulation the tests check it against).

Layer shapes follow the reference so that the parameter count (2,513,412) and a state-dict loaded
position-by-position from the reference model agree (tests/test_pwclite_structure.py).
`batch_directions=True` evaluates the two decoders (flows_12 / flows_21, :430-432) as ONE batch of
2B pairs: same result, half the launches, twice the work per kernel.
"""
import torch
import torch.nn.functional as F
from torch import nn

from .correlation import Correlation, cost_concat
from .segpool import segment_max_pool
from .warp import flow_warp

PYRAMID_CHANNELS = (3, 16, 32, 64, 96, 128, 192)  # models/pwclite.py:183
SEARCH_RANGE = 4
OUTPUT_LEVEL = 4


def conv_block(cin, cout, kernel_size=3, stride=1, dilation=1, relu=True):
    layers = [nn.Conv2d(cin, cout, kernel_size, stride, ((kernel_size - 1) * dilation) // 2, dilation, bias=True)]
    if relu:
        layers.append(nn.LeakyReLU(0.1, inplace=True))
    return nn.Sequential(*layers)


class FeaturePyramid(nn.Module):
    """Six stride-2 stages (models/pwclite.py:42-75); returns features coarsest first."""

    def __init__(self, channels=PYRAMID_CHANNELS):
        super().__init__()
        self.convs = nn.ModuleList(
            nn.Sequential(conv_block(ci, co, stride=2), conv_block(co, co)) for ci, co in zip(channels[:-1], channels[1:]))

    def forward(self, x):
        feats = [x]
        for stage in self.convs:
            x = stage(x)
            feats.append(x)
        return feats[::-1]


class FlowEstimator(nn.Module):
    """"Reduce" estimator (models/pwclite.py:100-120)."""

    def __init__(self, cin):
        super().__init__()
        self.conv1 = conv_block(cin, 128)
        self.conv2 = conv_block(128, 128)
        self.conv3 = conv_block(256, 96)
        self.conv4 = conv_block(224, 64)
        self.conv5 = conv_block(160, 32)
        self.feat_dim = 32
        self.predict_flow = conv_block(96, 2, relu=False)

    def forward(self, x):
        x1 = self.conv1(x)
        x2 = self.conv2(x1)
        x3 = self.conv3(torch.cat([x1, x2], 1))
        x4 = self.conv4(torch.cat([x2, x3], 1))
        x5 = self.conv5(torch.cat([x3, x4], 1))
        return x5, self.predict_flow(torch.cat([x4, x5], 1))


class ContextNetwork(nn.Module):
    """Dilated refinement (models/pwclite.py:123-139)."""

    def __init__(self, cin):
        super().__init__()
        self.convs = nn.Sequential(conv_block(cin, 128, 3, 1, 1), conv_block(128, 128, 3, 1, 2),
                                   conv_block(128, 128, 3, 1, 4), conv_block(128, 96, 3, 1, 8))
        self.flow_head = nn.Sequential(conv_block(96, 64, 3, 1, 16), conv_block(64, 32, 3, 1, 1),
                                       conv_block(32, 2, relu=False))

    def forward(self, x):
        feat = self.convs(x)
        return self.flow_head(feat), feat


class ConvexUpsampler(nn.Module):
    """Learned x4 convex upsampling (models/pwclite.py:142-166)."""

    def __init__(self, cin=96, scale=4):
        super().__init__()
        self.scale = scale
        self.convs = nn.Sequential(conv_block(cin, 128), conv_block(128, scale * scale * 9))

    def forward(self, flow, feat):
        s = self.scale
        n, _, h, w = flow.shape
        mask = torch.softmax((0.25 * self.convs(feat)).view(n, 1, 9, s, s, h, w), dim=2)
        patches = F.unfold(s * flow, [3, 3], padding=1).view(n, 2, 9, 1, 1, h, w)
        up = torch.sum(mask * patches, dim=2).permute(0, 1, 4, 2, 5, 3)
        return up.reshape(n, 2, s * h, s * w)


def segment_max_pool_torch(feat, seg):
    """PyTorch formulation of the pooling, kept as the checker of the CUDA kernel (tests) and for CPU structure
    tests; the model calls unsamflow_b200.segpool.segment_max_pool (csrc/segpool.cu).

    feat [B,C,h,w], seg [B,1,h,w] integer ids -> every pixel receives the maximum of `feat` over its segment.

    Restates models/pwclite.py:319-328: `amax(one_hot(seg) * feat[..., None], dim=(2, 3))` followed by the
    one-hot weighted sum.  In that product the pixels outside a segment contribute zeros, so the pooled value is
    max(0, max over the segment) unless the segment covers the whole image -- reproduced here by reducing into a
    zero-initialised table (and into -inf for an all-covering segment)."""
    b, c, h, w = feat.shape
    ids = seg.reshape(b, 1, h * w).long()
    n = int(ids.max().item()) + 1
    idx = ids.expand(b, c, h * w)
    table = feat.new_zeros(b, c, n)
    covers_all = feat.new_zeros(b, 1, n).scatter_add_(2, ids, feat.new_ones(b, 1, h * w)) == h * w
    table = table.masked_fill(covers_all, float("-inf"))
    table = table.scatter_reduce(2, idx, feat.reshape(b, c, h * w), "amax", include_self=True)
    return table.gather(2, idx).reshape(b, c, h, w)


class PWCLite(nn.Module):
    def __init__(self, batch_directions=False, add_mask_corr=False, aggregation_type="concat", num_segments=None,
                 fuse_cost_concat=False):
        """num_segments: one-hot classes of the segment maps (max id + 1); None reads it from the map on every call
        (a device->host sync, like the reference's F.one_hot); e.g. 256 for uint8 SAM id maps keeps the step sync-free."""
        super().__init__()
        self.num_segments = num_segments
        # True: the correlation kernels apply the LeakyReLU and write into the estimator's input buffer
        # (correlation.cost_concat) instead of corr -> LeakyReLU(inplace) -> torch.cat
        self.fuse_cost_concat = fuse_cost_concat
        self.batch_directions = batch_directions
        self.add_mask_corr = add_mask_corr
        self.aggregation_type = aggregation_type
        if aggregation_type not in ("concat", "residual"):
            raise NotImplementedError(f"aggregation_type={aggregation_type!r}")  # models/pwclite.py:350-351
        self.feature_pyramid_extractor = FeaturePyramid()
        self.corr = Correlation(pad_size=SEARCH_RANGE, kernel_size=1, max_displacement=SEARCH_RANGE, stride1=1,
                                stride2=1, corr_multiply=1)
        self.leaky = nn.LeakyReLU(0.1, inplace=True)
        dim_corr = (2 * SEARCH_RANGE + 1) ** 2
        self.flow_estimators = FlowEstimator(32 + (2 if add_mask_corr else 1) * dim_corr + 2)
        self.context_networks = ContextNetwork(self.flow_estimators.feat_dim + 2)
        self.output_flow_upsampler = ConvexUpsampler(96, 4)
        self.conv_1x1 = nn.ModuleList(conv_block(c, 32, kernel_size=1) for c in PYRAMID_CHANNELS[:1:-1])
        if add_mask_corr:  # models/pwclite.py:249-269 (registered after conv_1x1, like the reference)
            self.conv_1x1_mask = nn.ModuleList(conv_block(c, 32, kernel_size=1) for c in PYRAMID_CHANNELS[:1:-1])
            self.mask_aggregation = conv_block(64 if aggregation_type == "concat" else 32, 32, kernel_size=1)

    def mask_features(self, level, x, seg):
        """models/pwclite.py:317-351 for one image: 1x1 conv, pooling over the SAM segments, aggregation."""
        proj = self.conv_1x1_mask[level](x)
        pooled = segment_max_pool(proj, F.interpolate(seg, x.shape[-2:], mode="nearest"), self.num_segments)
        if self.aggregation_type == "residual":
            return proj + self.mask_aggregation(pooled)
        return self.mask_aggregation(torch.cat([proj, pooled], 1))

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
                nn.init.zeros_(m.bias)

    def num_parameters(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def decoder(self, pyr1, pyr2, seg1=None, seg2=None):
        flows = []
        b, _, h0, w0 = pyr1[0].shape
        flow = torch.zeros(b, 2, h0, w0, dtype=torch.float32, device=pyr1[0].device)
        for level, (x1, x2) in enumerate(zip(pyr1, pyr2)):
            if level > 0:
                flow = F.interpolate(flow * 2, scale_factor=2, mode="bilinear", align_corners=True)
                x2_warp = flow_warp(x2, flow)
            else:
                x2_warp = x2
            x1_proj = self.conv_1x1[level](x1)
            pairs = [(x1, x2_warp)]
            if self.add_mask_corr:
                m1, m2 = self.mask_features(level, x1, seg1), self.mask_features(level, x2, seg2)
                # the reference warps with `flow` at every level, the all-zero initial flow included (:355)
                pairs.append((m1, flow_warp(m2, flow)))
            if self.fuse_cost_concat:
                est_in = cost_concat(pairs, [x1_proj, flow], 0.1)
            else:
                est_in = torch.cat([self.leaky(self.corr(a, b)) for a, b in pairs] + [x1_proj, flow], 1)
            feat, flow_res = self.flow_estimators(est_in)
            flow = flow + flow_res
            flow_fine, up_feat = self.context_networks(torch.cat([feat, flow], 1))
            flow = flow + flow_fine
            flows.append(self.output_flow_upsampler(flow, up_feat))
            if level == OUTPUT_LEVEL:
                break
        return flows[::-1]

    def forward(self, img1, img2, full_seg1=None, full_seg2=None, with_bk=False):
        if self.add_mask_corr and (full_seg1 is None or full_seg2 is None):
            raise ValueError("add_mask_corr=True needs full_seg1 / full_seg2 ([B,1,H,W] segment ids)")
        if not self.add_mask_corr and (full_seg1 is not None or full_seg2 is not None):
            raise NotImplementedError("segmentation inputs without add_mask_corr (input_adj_map / input_boundary) "
                                      "are not restated here")
        feat1 = self.feature_pyramid_extractor(img1)
        feat2 = self.feature_pyramid_extractor(img2)
        if with_bk and self.batch_directions:
            n = img1.shape[0]
            seg12 = seg21 = None
            if self.add_mask_corr:
                seg12, seg21 = torch.cat([full_seg1, full_seg2]), torch.cat([full_seg2, full_seg1])
            both = self.decoder([torch.cat([a, b]) for a, b in zip(feat1, feat2)],
                                [torch.cat([b, a]) for a, b in zip(feat1, feat2)], seg12, seg21)
            return {"flows_12": [f[:n] for f in both], "flows_21": [f[n:] for f in both]}
        out = {"flows_12": self.decoder(feat1, feat2, full_seg1, full_seg2)}
        if with_bk:
            out["flows_21"] = self.decoder(feat2, feat1, full_seg2, full_seg1)
        return out
