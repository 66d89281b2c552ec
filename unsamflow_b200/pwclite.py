"""PWC-Lite with the B200-native hot path wired in: the CALLER of Correlation / flow_warp.

Scope note: the product of this repo is the correlation + warp path; this
module restates the network around it (models/pwclite.py:14-434, base configuration:
reduce_dense=True, learned_upsampler=True, no mask-correlation branch) only so that the path can
be exercised and timed inside a real training step without the reference tree on the box.  All
learned layers are plain PyTorch / cuDNN, exactly as in the reference; the two call sites are

    x2_warp = flow_warp(x2, flow)          models/pwclite.py:302
    out_corr = self.corr(x1, x2_warp)      models/pwclite.py:307

Layer shapes follow the reference so that the parameter count (2,513,412) and a state-dict loaded
position-by-position from the reference model agree (tests/test_pwclite_structure.py).
`batch_directions=True` evaluates the two decoders (flows_12 / flows_21, :430-432) as ONE batch of
2B pairs: same result, half the launches, twice the work per kernel.
"""
import torch
import torch.nn.functional as F
from torch import nn

from .correlation import Correlation
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


class PWCLite(nn.Module):
    def __init__(self, batch_directions=False):
        super().__init__()
        self.batch_directions = batch_directions
        self.feature_pyramid_extractor = FeaturePyramid()
        self.corr = Correlation(pad_size=SEARCH_RANGE, kernel_size=1, max_displacement=SEARCH_RANGE, stride1=1,
                                stride2=1, corr_multiply=1)
        self.leaky = nn.LeakyReLU(0.1, inplace=True)
        dim_corr = (2 * SEARCH_RANGE + 1) ** 2
        self.flow_estimators = FlowEstimator(32 + dim_corr + 2)
        self.context_networks = ContextNetwork(self.flow_estimators.feat_dim + 2)
        self.output_flow_upsampler = ConvexUpsampler(96, 4)
        self.conv_1x1 = nn.ModuleList(conv_block(c, 32, kernel_size=1) for c in PYRAMID_CHANNELS[:1:-1])

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
                nn.init.zeros_(m.bias)

    def num_parameters(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def decoder(self, pyr1, pyr2):
        flows = []
        b, _, h0, w0 = pyr1[0].shape
        flow = torch.zeros(b, 2, h0, w0, dtype=torch.float32, device=pyr1[0].device)
        for level, (x1, x2) in enumerate(zip(pyr1, pyr2)):
            if level > 0:
                flow = F.interpolate(flow * 2, scale_factor=2, mode="bilinear", align_corners=True)
                x2_warp = flow_warp(x2, flow)
            else:
                x2_warp = x2
            cost = self.leaky(self.corr(x1, x2_warp))
            feat, flow_res = self.flow_estimators(torch.cat([cost, self.conv_1x1[level](x1), flow], 1))
            flow = flow + flow_res
            flow_fine, up_feat = self.context_networks(torch.cat([feat, flow], 1))
            flow = flow + flow_fine
            flows.append(self.output_flow_upsampler(flow, up_feat))
            if level == OUTPUT_LEVEL:
                break
        return flows[::-1]

    def forward(self, img1, img2, full_seg1=None, full_seg2=None, with_bk=False):
        if full_seg1 is not None or full_seg2 is not None:
            raise NotImplementedError("segmentation inputs (add_mask_corr / input_adj_map) are not restated here")
        feat1 = self.feature_pyramid_extractor(img1)
        feat2 = self.feature_pyramid_extractor(img2)
        if with_bk and self.batch_directions:
            n = img1.shape[0]
            both = self.decoder([torch.cat([a, b]) for a, b in zip(feat1, feat2)],
                                [torch.cat([b, a]) for a, b in zip(feat1, feat2)])
            return {"flows_12": [f[:n] for f in both], "flows_21": [f[n:] for f in both]}
        out = {"flows_12": self.decoder(feat1, feat2)}
        if with_bk:
            out["flows_21"] = self.decoder(feat2, feat1)
        return out
