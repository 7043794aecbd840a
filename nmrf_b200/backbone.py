"""Feature extractor (NOT the hot path; kept in torch/cuDNN -- SURVEY.md §8(f) N2).

State-dict compatible with the reference's ResNet-style `Backbone` (nmrf/models/backbone.py:48-98):
7x7/2 stem, three stages of two residual blocks (64, 96/2, 128), InstanceNorm, 1x1 output conv;
returns [feat@1/4, feat@1/8 = avg_pool2(feat@1/4)].  A different encoder (e.g. the reference's
SwinAdaptor with `nmrf_b200.msda` as its MSDeformAttn op) can be passed to `NMRF(backbone=...)`
as long as it has `.output_dim` and the same return convention.
"""
import torch.nn as nn
import torch.nn.functional as F


class _Residual(nn.Module):
    def __init__(self, cin, cout, norm, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.relu = nn.ReLU(inplace=True)
        self.norm1, self.norm2 = norm(cout), norm(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.norm3 = norm(cout)
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride), self.norm3)

    def forward(self, x):
        y = self.relu(self.norm1(self.conv1(x)))
        y = self.relu(self.norm2(self.conv2(y)))
        return self.relu(y + (x if self.downsample is None else self.downsample(x)))


class Backbone(nn.Module):
    def __init__(self, output_dim=256, norm_layer=nn.InstanceNorm2d):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.norm1 = norm_layer(64)
        self.relu1 = nn.ReLU(inplace=True)
        self.layer1 = nn.Sequential(_Residual(64, 64, norm_layer, 1), _Residual(64, 64, norm_layer, 1))
        self.layer2 = nn.Sequential(_Residual(64, 96, norm_layer, 2), _Residual(96, 96, norm_layer, 1))
        self.layer3 = nn.Sequential(_Residual(96, 128, norm_layer, 1), _Residual(128, 128, norm_layer, 1))
        self.conv2 = nn.Conv2d(128, output_dim, 1)
        self.output_dim = output_dim
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x):
        x = 2 * (x / 255.0) - 1.0                       # backbone.py:86
        x = self.relu1(self.norm1(self.conv1(x)))
        x = self.conv2(self.layer3(self.layer2(self.layer1(x))))
        return [x, F.avg_pool2d(x, 2, 2)]
