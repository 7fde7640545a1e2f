"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Stand-in for the un-vendored third-party package ``resnest.torch`` that the
reference imports unconditionally (/root/reference/model/unet.py:4) and calls at
/root/reference/model/unet.py:52 as ``resnest[name](pretrained=..., dilation=...)``.

The reference pins no version (requirements.txt:1 is an unpinned git URL), so this
restates the published ResNeSt architecture (Zhang et al. 2020, "ResNeSt:
Split-Attention Networks"): deep-stem ResNet-D, radix-2 / cardinality-1
Split-Attention bottlenecks, avg-down shortcuts, avd pooling after the SplAt conv.
The restatement is pinned by the four published parameter counts, asserted in
tests/test_oracle.py (27 483 240 / 48 275 016 / 70 201 544 / 110 929 480).

Module / attribute names follow the upstream package because the reference
re-homes them into its own state_dict (unet.py:80-84) and indexes ``conv1[0]``
(unet.py:66); they are the checkpoint-key contract, not copied code.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

__all__ = ["resnest50", "resnest101", "resnest200", "resnest269", "ResNeSt"]


class RadixSoftmax(nn.Module):
    def __init__(self, radix, cardinality):
        super().__init__()
        self.radix, self.cardinality = radix, cardinality

    def forward(self, x):
        b = x.size(0)
        if self.radix == 1:
            return torch.sigmoid(x)
        x = x.view(b, self.cardinality, self.radix, -1).transpose(1, 2)
        return F.softmax(x, dim=1).reshape(b, -1)


class SplAtConv2d(nn.Module):
    """Split-attention conv: grouped kxk conv -> bn0 -> relu -> radix-sum -> GAP -> fc1 -> bn1 -> relu -> fc2 -> r-softmax."""

    def __init__(self, in_channels, channels, kernel_size, stride, padding, dilation, groups, radix, reduction_factor=4):
        super().__init__()
        inter = max(in_channels * radix // reduction_factor, 32)
        self.radix, self.cardinality, self.channels = radix, groups, channels
        self.conv = nn.Conv2d(in_channels, channels * radix, kernel_size, stride, padding, dilation, groups=groups * radix, bias=False)
        self.bn0 = nn.BatchNorm2d(channels * radix)
        self.relu = nn.ReLU(inplace=True)
        self.fc1 = nn.Conv2d(channels, inter, 1, groups=groups)
        self.bn1 = nn.BatchNorm2d(inter)
        self.fc2 = nn.Conv2d(inter, channels * radix, 1, groups=groups)
        self.rsoftmax = RadixSoftmax(radix, groups)

    def forward(self, x):
        x = self.relu(self.bn0(self.conv(x)))
        b, rc = x.shape[:2]
        parts = torch.split(x, rc // self.radix, dim=1) if self.radix > 1 else (x,)
        gap = sum(parts)
        gap = F.adaptive_avg_pool2d(gap, 1)
        gap = self.relu(self.bn1(self.fc1(gap)))
        att = self.rsoftmax(self.fc2(gap)).view(b, -1, 1, 1)
        if self.radix > 1:
            atts = torch.split(att, rc // self.radix, dim=1)
            out = sum(a * p for a, p in zip(atts, parts))
        else:
            out = att * x
        return out.contiguous()


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride, downsample, radix, cardinality, bottleneck_width, avd, dilation, is_first):
        super().__init__()
        gw = int(planes * (bottleneck_width / 64.0)) * cardinality
        self.conv1 = nn.Conv2d(inplanes, gw, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(gw)
        self.avd = avd and (stride > 1 or is_first)
        if self.avd:
            self.avd_layer = nn.AvgPool2d(3, stride, padding=1)
            stride = 1
        self.conv2 = SplAtConv2d(gw, gw, 3, stride, dilation, dilation, cardinality, radix)
        self.conv3 = nn.Conv2d(gw, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.conv2(out)
        if self.avd:
            out = self.avd_layer(out)
        out = self.bn3(self.conv3(out))
        res = x if self.downsample is None else self.downsample(x)
        out += res
        return self.relu(out)


class ResNeSt(nn.Module):
    def __init__(self, layers, stem_width, dilation=1, num_classes=1000, radix=2, cardinality=1, bottleneck_width=64):
        super().__init__()
        self.radix, self.cardinality, self.bottleneck_width = radix, cardinality, bottleneck_width
        sw = stem_width
        self.inplanes = sw * 2
        self.conv1 = nn.Sequential(
            nn.Conv2d(3, sw, 3, 2, 1, bias=False), nn.BatchNorm2d(sw), nn.ReLU(inplace=True),
            nn.Conv2d(sw, sw, 3, 1, 1, bias=False), nn.BatchNorm2d(sw), nn.ReLU(inplace=True),
            nn.Conv2d(sw, sw * 2, 3, 1, 1, bias=False),
        )
        self.bn1 = nn.BatchNorm2d(self.inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = self._stage(64, layers[0], 1, 1, is_first=False)
        self.layer2 = self._stage(128, layers[1], 2, 1)
        if dilation == 4:
            self.layer3 = self._stage(256, layers[2], 1, 2)
            self.layer4 = self._stage(512, layers[3], 1, 4)
        elif dilation == 2:
            self.layer3 = self._stage(256, layers[2], 2, 1)
            self.layer4 = self._stage(512, layers[3], 1, 2)
        else:
            self.layer3 = self._stage(256, layers[2], 2, 1)
            self.layer4 = self._stage(512, layers[3], 2, 1)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(512 * 4, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _stage(self, planes, blocks, stride, dilation, is_first=True):
        down = None
        if stride != 1 or self.inplanes != planes * 4:
            k = stride if dilation == 1 else 1
            down = nn.Sequential(
                nn.AvgPool2d(kernel_size=k, stride=k, ceil_mode=True, count_include_pad=False),
                nn.Conv2d(self.inplanes, planes * 4, 1, bias=False),
                nn.BatchNorm2d(planes * 4),
            )
        first_dil = 1 if dilation in (1, 2) else 2
        common = dict(radix=self.radix, cardinality=self.cardinality, bottleneck_width=self.bottleneck_width, avd=True)
        blks = [Bottleneck(self.inplanes, planes, stride, down, dilation=first_dil, is_first=is_first, **common)]
        self.inplanes = planes * 4
        for _ in range(1, blocks):
            blks.append(Bottleneck(self.inplanes, planes, 1, None, dilation=dilation, is_first=False, **common))
        return nn.Sequential(*blks)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(torch.flatten(self.avgpool(x), 1))


_SPECS = {"resnest50": ([3, 4, 6, 3], 32), "resnest101": ([3, 4, 23, 3], 64),
          "resnest200": ([3, 24, 36, 3], 64), "resnest269": ([3, 30, 48, 8], 64)}


def _factory(name):
    def build(pretrained=False, **kwargs):  # no network here: ``pretrained`` is accepted and ignored
        layers, sw = _SPECS[name]
        return ResNeSt(layers, sw, **kwargs)
    build.__name__ = name
    return build


resnest50, resnest101, resnest200, resnest269 = (_factory(n) for n in _SPECS)
