"""ResNeSt-50/101/200/269 and torchvision-style ResNet-50/101/152 encoders, already split into the five U-Net stages
the reference builds in get_encoder (/root/reference/model/unet.py:45-86).

The reference obtains these networks from un-vendored packages (resnest @ git HEAD, torchvision); their architecture is
restated here from the papers (ResNeSt: Zhang et al. 2020; ResNet v1.5) with module names chosen so that the state_dict
keys equal the reference's after its re-homing into nn.Sequential stages:
    enc_l1.0.{0,1,3,4,6}.* (ResNeSt deep stem) | enc_l1.0.* (ResNet 7x7)   enc_l1.1.* (bn1)
    enc_l2.1.<blk>.*  (index 0 is the parameter-free max-pool)             enc_l3..5.<blk>.*
All arithmetic runs through xview2_b200.ops.
"""
import math

import torch
from torch import nn

from .. import ops
from ..lib import ACT_NONE, ACT_RELU
from .layers import _cl, conv_param, run_conv

RESNEST_SPECS = {"resnest50": ([3, 4, 6, 3], 32), "resnest101": ([3, 4, 23, 3], 64),
                 "resnest200": ([3, 24, 36, 3], 64), "resnest269": ([3, 30, 48, 8], 64)}
RESNET_SPECS = {"resnet50": [3, 4, 6, 3], "resnet101": [3, 4, 23, 3], "resnet152": [3, 8, 36, 3]}


def _named(parent, pairs):
    for name, mod in pairs:
        parent.add_module(name, mod)
    return parent


def _sub(module, name):
    return module._modules[name]


# ---------------------------------------------------------------------------------------------------------------
# ResNeSt
# ---------------------------------------------------------------------------------------------------------------
class SplAtConv2d(nn.Module):
    """Radix-2 / cardinality-1 split-attention conv.  conv -> bn0 -> relu -> split attention (one fused tape node)."""

    def __init__(self, channels, dilation):
        super().__init__()
        inter = max(channels * 2 // 4, 32)
        self.conv = conv_param(channels, channels * 2, 3, 1, dilation, dilation, groups=2)
        self.bn0 = nn.BatchNorm2d(channels * 2)
        self.fc1 = nn.Conv2d(channels, inter, 1)
        self.bn1 = nn.BatchNorm2d(inter)
        self.fc2 = nn.Conv2d(inter, channels * 2, 1)

    def forward(self, x):
        return ops.conv_bn_split_attention(x, self.conv, self.bn0, self.fc1, self.bn1, self.fc2)


class SplAtBottleneck(nn.Module):
    def __init__(self, inplanes, planes, stride, dilation, is_first, with_down, down_pool):
        super().__init__()
        self.conv1 = conv_param(inplanes, planes, 1)
        self.bn1 = nn.BatchNorm2d(planes)
        self.avd_stride = stride if (stride > 1 or is_first) else 0  # avd pool after the SplAt conv
        self.conv2 = SplAtConv2d(planes, dilation)
        self.conv3 = conv_param(planes, planes * 4, 1)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.down_pool = down_pool
        if with_down:  # avg-down shortcut: index 0 is the parameter-free AvgPool2d
            self.downsample = _named(nn.Module(), [("1", conv_param(inplanes, planes * 4, 1)),
                                                   ("2", nn.BatchNorm2d(planes * 4))])
        else:
            self.downsample = None

    def forward(self, x):
        # identity shortcut: x feeds conv1 and the shortcut only -- their two gradients meet inside the producer's BN backward
        x, res = ops.fork(x) if self.downsample is None else (x, x)
        out = ops.conv_bn_act(x, self.conv1, self.bn1, ACT_RELU)
        out = self.conv2(out)
        if self.avd_stride:
            out = ops.avg_pool2d(out, 3, self.avd_stride, 1)
        if self.downsample is not None:
            if self.down_pool > 1:
                res = ops.avg_pool2d(res, self.down_pool, self.down_pool, 0, ceil_mode=True, count_include_pad=False)
            res = ops.conv_bn_act(res, _sub(self.downsample, "1"), _sub(self.downsample, "2"), ACT_NONE)
        # bn3 + residual add + relu in one apply pass
        return ops.conv_bn_act(out, self.conv3, self.bn3, ACT_RELU, residual=res)


class _BlockList(nn.Module):
    """Children named "0", "1", ... executed in order (mirrors nn.Sequential naming without its forward)."""

    def __init__(self, blocks):
        super().__init__()
        for i, b in enumerate(blocks):
            self.add_module(str(i), b)

    def forward(self, x):
        for b in self._modules.values():
            x = b(x)
        return x


class ResNeStStem(nn.Module):
    """enc_l1: Sequential(conv1 = Sequential(conv, bn, relu, conv, bn, relu, conv), bn1, relu)."""

    def __init__(self, stem_width, in_channels=3):
        super().__init__()
        sw = stem_width
        stem = _named(nn.Module(), [("0", conv_param(in_channels, sw, 3, 2, 1)), ("1", nn.BatchNorm2d(sw)),
                                    ("3", conv_param(sw, sw, 3, 1, 1)), ("4", nn.BatchNorm2d(sw)),
                                    ("6", conv_param(sw, sw * 2, 3, 1, 1))])
        self.add_module("0", stem)
        self.add_module("1", nn.BatchNorm2d(sw * 2))

    def forward(self, x):
        s = _sub(self, "0")
        x = ops.conv_bn_act(x, _sub(s, "0"), _sub(s, "1"), ACT_RELU)
        x = ops.conv_bn_act(x, _sub(s, "3"), _sub(s, "4"), ACT_RELU)
        return ops.conv_bn_act(x, _sub(s, "6"), _sub(self, "1"), ACT_RELU)


class PooledStage(nn.Module):
    """enc_l2: Sequential(maxpool, layer1) -- the blocks live under child "1"."""

    def __init__(self, blocks):
        super().__init__()
        self.add_module("1", _BlockList(blocks))

    def forward(self, x):
        return _sub(self, "1")(ops.max_pool2d(x, 3, 2, 1))


def _stage_plan(dilation):
    if dilation == 1:
        return [(1, 1, 1), (2, 1, 1), (2, 1, 1), (2, 1, 1)]
    if dilation == 2:
        return [(1, 1, 1), (2, 1, 1), (2, 1, 1), (1, 1, 2)]
    if dilation == 4:
        return [(1, 1, 1), (2, 1, 1), (1, 1, 2), (1, 2, 4)]
    raise ValueError("Dilation can be set to 1, 2 or 4")


def _init_resnest(module):
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
            m.weight.data.normal_(0, math.sqrt(2.0 / n))
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data.fill_(1)
            m.bias.data.zero_()


def build_resnest(name, dilation, in_channels=3):
    layers, sw = RESNEST_SPECS[name]
    inplanes = sw * 2
    stages = []
    for i, (planes, (stride, d_first, d_rest)) in enumerate(zip((64, 128, 256, 512), _stage_plan(dilation))):
        blocks = [SplAtBottleneck(inplanes, planes, stride, d_first, is_first=(i != 0), with_down=True, down_pool=stride)]
        inplanes = planes * 4
        blocks += [SplAtBottleneck(inplanes, planes, 1, d_rest, False, False, 1) for _ in range(1, layers[i])]
        stages.append(blocks)
    l1 = ResNeStStem(sw, in_channels)
    l2 = PooledStage(stages[0])
    l3, l4, l5 = (_BlockList(b) for b in stages[1:])
    for m in (l1, l2, l3, l4, l5):
        _init_resnest(m)
    return [2 * sw, 256, 512, 1024, 2048], l1, l2, l3, l4, l5


# ---------------------------------------------------------------------------------------------------------------
# ResNet (torchvision layout, v1.5: stride on the 3x3)
# ---------------------------------------------------------------------------------------------------------------
class Bottleneck(nn.Module):
    def __init__(self, inplanes, planes, stride, dilation, with_down):
        super().__init__()
        self.conv1 = conv_param(inplanes, planes, 1)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = conv_param(planes, planes, 3, stride, dilation, dilation)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = conv_param(planes, planes * 4, 1)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        if with_down:
            self.downsample = _named(nn.Module(), [("0", conv_param(inplanes, planes * 4, 1, stride)),
                                                   ("1", nn.BatchNorm2d(planes * 4))])
        else:
            self.downsample = None

    def forward(self, x):
        x, res = ops.fork(x) if self.downsample is None else (x, x)
        out = ops.conv_bn_act(x, self.conv1, self.bn1, ACT_RELU)
        out = ops.conv_bn_act(out, self.conv2, self.bn2, ACT_RELU)
        if self.downsample is not None:
            res = ops.conv_bn_act(x, _sub(self.downsample, "0"), _sub(self.downsample, "1"), ACT_NONE)
        return ops.conv_bn_act(out, self.conv3, self.bn3, ACT_RELU, residual=res)


class ResNetStem(nn.Module):
    """enc_l1: Sequential(conv1 7x7 s2, bn1, relu)."""

    def __init__(self, in_channels=3):
        super().__init__()
        self.add_module("0", conv_param(in_channels, 64, 7, 2, 3, bias=(in_channels != 3)))  # unet.py:67-74 passes bias=conv1.bias
        self.add_module("1", nn.BatchNorm2d(64))

    def forward(self, x):
        return ops.conv_bn_act(x, _sub(self, "0"), _sub(self, "1"), ACT_RELU)


def build_resnet(name, dilation, in_channels=3):
    layers = RESNET_SPECS[name]
    inplanes = 64
    stages = []
    for i, (planes, (stride, d_first, d_rest)) in enumerate(zip((64, 128, 256, 512), _stage_plan(dilation))):
        blocks = [Bottleneck(inplanes, planes, stride, d_first, True)]
        inplanes = planes * 4
        blocks += [Bottleneck(inplanes, planes, 1, d_rest, False) for _ in range(1, layers[i])]
        stages.append(blocks)
    l1 = ResNetStem(3)
    if in_channels != 3:
        l1 = ResNetStem(in_channels)
    l2 = PooledStage(stages[0])
    l3, l4, l5 = (_BlockList(b) for b in stages[1:])
    for mod in (l1, l2, l3, l4, l5):  # torchvision init: kaiming_normal_(fan_out, relu) for convs, BN (1, 0)
        for m in mod.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                _cl(m)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
    return [64, 256, 512, 1024, 2048], l1, l2, l3, l4, l5
