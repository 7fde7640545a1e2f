"""Decoder / head building blocks with the reference's class names, constructor signatures and parameter names
(/root/reference/model/layers.py) so checkpoints load strictly -- but every forward pass runs libxv2 kernels.

The torch.nn layers created here are PARAMETER CONTAINERS only (they define state_dict keys and initialisation);
their own forward() is never called.
"""
import torch
from torch import nn

from .. import ops
from ..lib import ACT_LRELU, ACT_NONE, ACT_RELU


def _cl(module):
    """Keeps 4-D weights physically channels-last ([K][R][S][C]) so kernels and gradients share one layout."""
    for p in module.parameters(recurse=False):
        if p.dim() == 4:
            p.data = p.data.contiguous(memory_format=torch.channels_last)
    return module


def conv_param(cin, cout, k, stride=1, padding=0, dilation=1, groups=1, bias=False):
    return _cl(nn.Conv2d(cin, cout, k, stride, padding, dilation, groups, bias))


def run_conv(conv, x, x2=None):
    """Runs the nn.Conv2d parameter container `conv` through the xv2 convolution (optionally on cat(x, x2))."""
    return ops.conv2d(x, conv.weight, conv.bias, conv.stride[0], conv.padding[0], conv.dilation[0], conv.groups, x2)


class ConvLayer(nn.Module):
    """3x3 conv (no bias) -> BatchNorm -> LeakyReLU(0.01).  layers.py:89-100"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = conv_param(in_channels, out_channels, 3, padding=1)
        self.batch_norm = nn.BatchNorm2d(out_channels, affine=True)
        self.lrelu = nn.LeakyReLU(negative_slope=0.01, inplace=True)  # kept for module-tree parity (no parameters)

    def forward(self, inputs, second=None, defer=False):
        """`second`: optional tensor concatenated after `inputs` on channels without materialising the cat.
        `defer`: return the raw conv output with its BatchNorm + LeakyReLU pending (ops.DeferredBNAct) so that the output head
        can apply them in its own pass (the full-resolution activation is then never written)."""
        if defer:
            return ops.conv_bn_act_deferred(inputs, self.conv, self.batch_norm, ACT_LRELU, x2=second)
        return ops.conv_bn_act(inputs, self.conv, self.batch_norm, ACT_LRELU, x2=second)


class ConvBlock(nn.Module):
    """layers.py:119-128"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv1 = ConvLayer(in_channels, out_channels)
        self.conv2 = ConvLayer(out_channels, out_channels)

    def forward(self, inputs, second=None, defer=False):
        return self.conv2(self.conv1(inputs, second), defer=defer)


class AttentionLayer(nn.Module):
    """1x1 conv (no bias) -> BatchNorm.  layers.py:68-77"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = conv_param(in_channels, out_channels, 1)
        self.batch_norm = nn.BatchNorm2d(out_channels, affine=True)

    def forward(self, inputs, act=ACT_NONE, residual=None):
        return ops.conv_bn_act(inputs, self.conv, self.batch_norm, act, residual=residual)


class ConvTranspose(nn.Module):
    """2x2 stride-2 transposed conv, no bias.  layers.py:80-86"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = _cl(nn.ConvTranspose2d(in_channels, out_channels, kernel_size=2, stride=2, bias=False))

    def forward(self, inputs):
        return ops.conv_transpose2x2(inputs, self.conv.weight)


class UpsampleBlock(nn.Module):
    """Decoder stage: up-sample -> [attention gate] -> cat(skip) -> ConvBlock.  layers.py:131-168

    The cat is never materialised: the first 3x3 conv reads (up, skip) as two sources.
    """

    def __init__(self, in_channels, out_channels, skip_channels, attention, dec_interp):
        super().__init__()
        self.attention = attention
        self.dec_interp = dec_interp
        self.skip_channels = skip_channels
        if dec_interp:  # layers.py:138-139: 3x3 conv WITH bias, then bilinear x2 (align_corners=True)
            self.conv = conv_param(in_channels, out_channels, 3, padding=1, bias=True)
        else:
            self.conv_tranpose = ConvTranspose(in_channels, out_channels)  # (sic) the reference's attribute name
        self.conv_block = ConvBlock(skip_channels + out_channels, out_channels)
        if skip_channels > 0 and attention:
            att = out_channels // 2
            self.conv_o = AttentionLayer(out_channels, att)
            self.conv_s = AttentionLayer(skip_channels, att)
            self.psi = AttentionLayer(att, 1)
            self.sigmoid = nn.Sigmoid()
            self.relu = nn.ReLU(inplace=True)

    def forward(self, inputs, skip, defer=False):
        if self.dec_interp:
            out = run_conv(self.conv, inputs)
            out = ops.bilinear(out, (2 * out.shape[2], 2 * out.shape[3]))
        else:
            out = self.conv_tranpose(inputs)
        if self.skip_channels == 0:
            return self.conv_block(out, defer=defer)
        if self.attention:
            # relu(conv_o(out) + conv_s(skip)): the add + relu ride on conv_s's BN apply pass
            mix = self.conv_s(skip, ACT_RELU, residual=self.conv_o(out))
            skip = ops.gate(skip, self.psi(mix))
        return self.conv_block(out, skip, defer=defer)


class FusionBlock(nn.Module):
    """layers.py:103-116: run the pre/post stage, then two ConvLayer(2C -> C) on cat(pre, post) (cat not materialised)."""

    def __init__(self, pre_conv, post_conv, channels):
        super().__init__()
        self.pre_conv = pre_conv
        self.post_conv = post_conv
        self.conv_pre = ConvLayer(2 * channels, channels)
        self.conv_post = ConvLayer(2 * channels, channels)

    def forward(self, pre, post, dec_pre=None, dec_post=None, last_dec=False):
        pre = self.pre_conv(pre, dec_pre) if dec_pre is not None or last_dec else self.pre_conv(pre)
        post = self.post_conv(post, dec_post) if dec_post is not None or last_dec else self.post_conv(post)
        return self.conv_pre(pre, post), self.conv_post(pre, post)


class OutputBlock(nn.Module):
    """1x1 conv + bias to n_class logits (fp32); CORAL head: 1-channel conv + three rank biases; --interpolate: bilinear
    resize of the logits to 512^2 (training) / 1024^2 (eval).  layers.py:171-189"""

    def __init__(self, in_channels, nclass, interpolate):
        super().__init__()
        self.interpolate = interpolate
        self.coral_loss = nclass == 3
        if self.coral_loss:
            self.conv = nn.Conv2d(in_channels, 1, kernel_size=1, bias=False)
            self.bias = nn.Parameter(torch.tensor([[[1.0]], [[0.0]], [[-1.0]]]))
        else:
            self.conv = nn.Conv2d(in_channels, nclass, kernel_size=1)

    def forward(self, inputs, second=None):
        if isinstance(inputs, ops.DeferredBNAct):
            if second is None and not self.coral_loss and not self.interpolate:
                return ops.bnact_head(inputs, self.conv.weight, self.conv.bias)
            inputs = inputs.materialise()
        if isinstance(second, ops.DeferredBNAct):
            second = second.materialise()
        if second is not None:  # Siamese / fused heads read cat(pre, post): tiny-N GEMM, concatenate then stream once
            inputs = torch.cat((inputs, second), 1)
        if self.coral_loss:  # one shared projection + per-rank bias == a 3-row head whose rows alias the same weights
            weight, bias = self.conv.weight.expand(3, -1, -1, -1), self.bias.reshape(3)
        else:
            weight, bias = self.conv.weight, self.conv.bias
        if self.interpolate:
            # the head sits on the last ENCODER stage here (2048 / 4096 channels at 1/32 resolution): a general fp32 conv,
            # then the bilinear resize of the logits (layers.py:186-188)
            out = ops.conv2d(ops.cast(inputs, torch.float32), weight.contiguous(), bias)
            size = (512, 512) if self.training else (1024, 1024)
            return ops.bilinear(out, size)
        return ops.head(inputs, weight, bias)


class _PPMBranch(nn.Module):
    """Children "1" (1x1 conv, no bias) and "2" (BatchNorm): the parameterised members of the reference's
    nn.Sequential(AdaptiveAvgPool2d, Conv2d, BatchNorm2d, LeakyReLU) (layers.py:12-19), so the keys are features.<i>.1/2.*"""

    def __init__(self, in_channels, out_channels, bins):
        super().__init__()
        self.bins = bins
        self.add_module("1", conv_param(in_channels, out_channels, 1))
        self.add_module("2", nn.BatchNorm2d(out_channels, affine=True))

    def forward(self, x):
        pooled = ops.adaptive_avg_pool2d(x, self.bins)
        return ops.conv_bn_act(pooled, self._modules["1"], self._modules["2"], ACT_LRELU)


class PPM(nn.Module):
    """Pyramid pooling (layers.py:6-29): bins 1, 2, 3, 6 -> 1x1 conv -> BN -> LeakyReLU -> bilinear back to the input size,
    concatenated with the input, 1x1 conv + bias back to in_channels."""

    def __init__(self, in_channels):
        super().__init__()
        out_channels = in_channels // 4
        self.features = nn.ModuleList([_PPMBranch(in_channels, out_channels, b) for b in (1, 2, 3, 6)])
        self.conv = conv_param(2 * in_channels, in_channels, 1, bias=True)

    def forward(self, x):
        size = x.shape[2:]
        outs = [x] + [ops.bilinear(f(x), size) for f in self.features]
        return run_conv(self.conv, torch.cat(outs, 1))  # the cat is a 1/32-resolution plumbing copy


class ASPPModule(nn.Module):
    """conv (1x1 or dilated 3x3, no bias) -> BN -> LeakyReLU.  layers.py:32-48"""

    def __init__(self, in_channels, out_channels, kernel_size, padding, dilation):
        super().__init__()
        self.conv = conv_param(in_channels, out_channels, kernel_size, 1, padding, dilation)
        self.bn = nn.BatchNorm2d(out_channels, affine=True)
        self.relu = nn.LeakyReLU(negative_slope=0.01, inplace=True)
        torch.nn.init.kaiming_normal_(self.conv.weight)
        _cl(self.conv)

    def forward(self, x):
        return ops.conv_bn_act(x, self.conv, self.bn, ACT_LRELU)


class ASPP(nn.Module):
    """Atrous spatial pyramid (layers.py:51-65): dilations 1, 3d, 6d, 9d, branches concatenated (4 x C/4 = C channels)."""

    def __init__(self, in_channels, dilation):
        super().__init__()
        out_channels = in_channels // 4
        d = [1, 3 * dilation, 6 * dilation, 9 * dilation]
        self.aspp1 = ASPPModule(in_channels, out_channels, 1, padding=0, dilation=d[0])
        self.aspp2 = ASPPModule(in_channels, out_channels, 3, padding=d[1], dilation=d[1])
        self.aspp3 = ASPPModule(in_channels, out_channels, 3, padding=d[2], dilation=d[2])
        self.aspp4 = ASPPModule(in_channels, out_channels, 3, padding=d[3], dilation=d[3])

    def forward(self, x):
        return torch.cat((self.aspp1(x), self.aspp2(x), self.aspp3(x), self.aspp4(x)), dim=1)
