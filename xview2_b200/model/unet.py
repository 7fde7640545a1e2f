"""U-Net assembly with the reference's construction API (/root/reference/model/unet.py): get_encoder, get_decoder,
UNetTemplate, OutputTemplate, UNetLoc, get_dmg_unet and the eight damage-model classes -- same names, signatures,
attribute names (hence state_dict keys) and the same quirks where they affect results (see README "Reference quirks").
Forward passes run on libxv2 kernels; the compute dtype comes from ``args.precision`` (bf16 tensor-core path unless 32).
"""
import torch
from torch import nn

from .. import ops
from .encoders import RESNEST_SPECS, RESNET_SPECS, build_resnest, build_resnet
from .layers import ASPP, PPM, FusionBlock, OutputBlock, UpsampleBlock

DEC_CHANNELS = [512, 256, 128, 64, 32]


def compute_dtype(args):
    """--precision 32 -> fp32 parity path; 16 (reference AMP default, main.py:36) and 'bf16' -> bf16 tensor cores."""
    return torch.float32 if str(getattr(args, "precision", "bf16")) == "32" else torch.bfloat16


def concat(x, y):
    return None if x is None or y is None else torch.cat([x, y], 1)


def get_nclass(args):
    if args.loss_str == "mse":
        return 1
    if args.loss_str == "coral":
        return 3
    return 4


def get_encoder(encoder_str, dilation, pretrained=True, in_channels=3):
    """Returns (channels, l1, l2, l3, l4, l5).  `pretrained` is accepted for signature parity; weights are seeded-random
    here (no network) and are normally overwritten by a checkpoint."""
    assert "resnet" in encoder_str or "resnest" in encoder_str
    if encoder_str in RESNEST_SPECS:
        return build_resnest(encoder_str, dilation, in_channels)
    if encoder_str in RESNET_SPECS:
        return build_resnet(encoder_str, dilation, in_channels)
    raise ValueError(f"Not implemented encoder {encoder_str}")


def get_decoder(encf, dilation, attn, no_skip=False, dec_interp=False):
    """Five decoder stages; dilation 2 / 4 drop the first one / two.  unet.py:89-110"""
    if dilation not in (1, 2, 4):
        raise ValueError("Dilation can be set to 1, 2 or 4")
    first = {1: 0, 2: 1, 4: 2}[dilation]
    stages = [None] * 5
    cin = encf[-1]
    for i in range(first, 5):
        skip = 0 if (no_skip or i == 4) else encf[-2 - i]
        stages[i] = UpsampleBlock(cin, DEC_CHANNELS[i], skip, attn, dec_interp)
        cin = DEC_CHANNELS[i]
    return (DEC_CHANNELS, *stages)


def _decode(stages, encs, dilation, no_skip, defer_tail=False):
    """Shared decoder walk (unet.py:153-170): stages = [dec_l1..dec_l5] (None when dropped), encs = [enc1..enc5].
    `defer_tail`: dec5 feeds nothing but the output head, so its last BatchNorm + LeakyReLU may stay pending for the head."""
    first = {1: 0, 2: 1, 4: 2}[dilation]
    x = encs[4]
    outs = {}
    for i in range(first, 5):
        skip = None if (no_skip or i == 4) else encs[3 - i]
        x = stages[i](x, skip, defer=True) if (defer_tail and i == 4) else stages[i](x, skip)
        outs[i] = x
    return outs[4], outs[3], outs[2]


class _Net(nn.Module):
    """Entry-point mix-in: casts the (fp32, NCHW) batch to the compute dtype and channels-last once."""

    def _prep(self, data):
        return ops.cast(data, self.compute_dtype)


class UNetTemplate(nn.Module):
    def __init__(self, args, in_channels=3):
        super().__init__()
        self.use_ppm = args.ppm
        self.use_aspp = args.aspp
        self.dilation = args.dilation
        self.no_skip = args.no_skip
        self.interpolate = args.interpolate
        self.enc_chn, self.enc_l1, self.enc_l2, self.enc_l3, self.enc_l4, self.enc_l5 = get_encoder(
            args.encoder, self.dilation, in_channels=in_channels)
        if self.use_ppm:
            self.ppm = PPM(self.enc_chn[-1])
        elif self.use_aspp:
            self.aspp = ASPP(self.enc_chn[-1], self.dilation)
        self.dec_chn = None
        if not self.interpolate:
            self.dec_chn, self.dec_l1, self.dec_l2, self.dec_l3, self.dec_l4, self.dec_l5 = get_decoder(
                self.enc_chn, self.dilation, args.attention, self.no_skip, args.dec_interp)

    def encode(self, data):
        # every stage output feeds the next stage AND the decoder (skip): ops.fork keeps autograd from summing the two gradient
        # parts in a pass of its own -- the stage's last BatchNorm backward adds them while it streams (s*: the skip aliases)
        enc1, s1 = ops.fork(self.enc_l1(data))
        enc2, s2 = ops.fork(self.enc_l2(enc1))
        enc3, s3 = ops.fork(self.enc_l3(enc2))
        enc4, s4 = ops.fork(self.enc_l4(enc3))
        enc5 = self.enc_l5(enc4)
        if self.use_ppm:      # unet.py:144-147
            enc5 = self.ppm(enc5)
        elif self.use_aspp:
            enc5 = self.aspp(enc5)
        return [s1, s2, s3, s4, enc5]

    def forward(self, data, defer_tail=False):
        encs = self.encode(data)
        if self.interpolate:  # unet.py:148-149: the head works on the last encoder stage
            return encs[4], None, None
        stages = [self.dec_l1, self.dec_l2, self.dec_l3, self.dec_l4, self.dec_l5]
        return _decode(stages, encs, self.dilation, self.no_skip, defer_tail)


class OutputTemplate(nn.Module):
    def __init__(self, n_class, deep_supervision, dec_chn, scale=1, interp=False, enc_last=0):
        super().__init__()
        self.deep_supervision = deep_supervision
        self.interp = interp
        if self.interp:  # unet.py:180-182
            d3 = d4 = None
            d5 = enc_last * scale
            self.deep_supervision = False
        else:
            d3, d4, d5 = scale * dec_chn[-3], scale * dec_chn[-2], scale * dec_chn[-1]
        if self.deep_supervision:
            self.output_block_ds3 = OutputBlock(d3, n_class, interp)
            self.output_block_ds4 = OutputBlock(d4, n_class, interp)
        self.output_block = OutputBlock(d5, n_class, interp)

    def forward(self, dec5, dec4, dec3, dec5b=None, dec4b=None, dec3b=None):
        """The optional *b tensors are the second half of a channel concat (Siamese / fused heads)."""
        out = self.output_block(dec5, dec5b)
        if self.training and self.deep_supervision:
            return [out, self.output_block_ds4(dec4, dec4b), self.output_block_ds3(dec3, dec3b)]
        return out


class UNetLoc(_Net):
    def __init__(self, args, in_channels=3, n_class=2):
        super().__init__()
        self.compute_dtype = compute_dtype(args)
        self.unet = UNetTemplate(args, in_channels)
        self.output_block = OutputTemplate(n_class, args.deep_supervision, self.unet.dec_chn, interp=args.interpolate,
                                           enc_last=self.unet.enc_chn[-1])

    def forward(self, data):
        # dec5 is read by the output head only: its BatchNorm + LeakyReLU are applied inside the head's own pass
        return self.output_block(*self.unet(self._prep(data), defer_tail=True))


class SiameseUNet(_Net):
    """One shared U-Net run on the pre then the post image (separate BN statistics, unet.py:231-233), heads on the cat."""

    def __init__(self, args, n_class):
        super().__init__()
        self.compute_dtype = compute_dtype(args)
        self.unet = UNetTemplate(args)
        self.output_block = OutputTemplate(n_class, args.deep_supervision, self.unet.dec_chn, 2, args.interpolate,
                                           self.unet.enc_chn[-1])

    def forward(self, data):
        data = self._prep(data)
        pre = self.unet(data[:, :3])
        post = self.unet(data[:, 3:])
        return self.output_block(*pre, *post)


class _TwinEncoderMixin:
    def _decode_cat(self, encs_pre, encs_post):
        encs = [concat(a, b) for a, b in zip(encs_pre, encs_post)]
        stages = [self.dec_l1, self.dec_l2, self.dec_l3, self.dec_l4, self.dec_l5]
        return _decode(stages, encs, self.dilation, self.no_skip)


class SiameseEncUNet(_Net, _TwinEncoderMixin):
    def __init__(self, args, n_class):
        super().__init__()
        self.compute_dtype = compute_dtype(args)
        self.use_ppm, self.use_aspp = args.ppm, args.aspp
        self.dilation, self.no_skip = args.dilation, args.no_skip
        if args.loss_str == "mse":
            n_class = 1
        elif args.loss_str == "level":
            n_class = 4
        self.enc_chn, self.enc_l1, self.enc_l2, self.enc_l3, self.enc_l4, self.enc_l5 = get_encoder(args.encoder, self.dilation)
        if self.use_ppm:
            self.ppm = PPM(self.enc_chn[-1])
        elif self.use_aspp:
            self.aspp = ASPP(self.enc_chn[-1], self.dilation)
        self.enc_chn = [2 * c for c in self.enc_chn]
        self.dec_chn, self.dec_l1, self.dec_l2, self.dec_l3, self.dec_l4, self.dec_l5 = get_decoder(
            self.enc_chn, self.dilation, args.attention, self.no_skip, args.dec_interp)
        self.output_block = OutputTemplate(n_class, args.deep_supervision, self.dec_chn, 1)

    def forward_enc(self, data):
        e1 = self.enc_l1(data)
        e2 = self.enc_l2(e1)
        e3 = self.enc_l3(e2)
        e4 = self.enc_l4(e3)
        e5 = self.enc_l5(e4)
        if self.use_ppm:
            e5 = self.ppm(e5)
        elif self.use_aspp:
            e5 = self.aspp(e5)
        return [e1, e2, e3, e4, e5]

    def forward(self, data):
        data = self._prep(data)
        return self.output_block(*self._decode_cat(self.forward_enc(data[:, :3]), self.forward_enc(data[:, 3:])))


class _FusedEncoders(_Net):
    """Twin encoders with a FusionBlock after every stage (unet.py:323-337).  Each stage is registered twice, as
    ``enc_lN_pre`` and as ``fusion_blockN.pre_conv`` -- the reference's state_dict has both spellings."""

    def _build_encoders(self, args):
        self.use_ppm, self.use_aspp = args.ppm, args.aspp
        self.dilation = 1
        _, self.enc_l1_pre, self.enc_l2_pre, self.enc_l3_pre, self.enc_l4_pre, self.enc_l5_pre = get_encoder(
            args.encoder, self.dilation, in_channels=3)
        enc_chn, self.enc_l1_post, self.enc_l2_post, self.enc_l3_post, self.enc_l4_post, self.enc_l5_post = get_encoder(
            args.encoder, self.dilation, in_channels=3)
        for i in range(1, 6):
            blk = FusionBlock(getattr(self, f"enc_l{i}_pre"), getattr(self, f"enc_l{i}_post"), enc_chn[i - 1])
            setattr(self, f"fusion_block{i}", blk)
        return enc_chn

    def _encode(self, data):
        pre, post = data[:, :3], data[:, 3:]
        pres, posts = [], []
        for i in range(1, 6):
            pre, post = getattr(self, f"fusion_block{i}")(pre, post)
            pres.append(pre)
            posts.append(post)
        return pres, posts


class FusedUNet(_FusedEncoders):
    def __init__(self, args, n_class):
        super().__init__()
        self.compute_dtype = compute_dtype(args)
        enc_chn = self._build_encoders(args)
        # the reference passes args.dec_interp in get_decoder's no_skip slot (unet.py:339-345); dec_interp is
        # unsupported here, so no_skip is always False exactly as in every reference run that works
        _, self.dec_l1_pre, self.dec_l2_pre, self.dec_l3_pre, self.dec_l4_pre, self.dec_l5_pre = get_decoder(
            enc_chn, self.dilation, args.attention, args.dec_interp)
        dec_chn, self.dec_l1_post, self.dec_l2_post, self.dec_l3_post, self.dec_l4_post, self.dec_l5_post = get_decoder(
            enc_chn, self.dilation, args.attention, args.dec_interp)
        for i in range(1, 6):
            blk = FusionBlock(getattr(self, f"dec_l{i}_pre"), getattr(self, f"dec_l{i}_post"), dec_chn[i - 1])
            setattr(self, f"fusion_block_dec{i}", blk)
        self.output_block = OutputTemplate(n_class, args.deep_supervision, dec_chn, 2)

    def forward(self, data):
        pres, posts = self._encode(self._prep(data))
        pre, post = pres[4], posts[4]
        dpre, dpost = [], []
        for i in range(1, 6):
            blk = getattr(self, f"fusion_block_dec{i}")
            if i < 5:
                pre, post = blk(pre, post, pres[4 - i], posts[4 - i])
            else:
                pre, post = blk(pre, post, last_dec=True)
            dpre.append(pre)
            dpost.append(post)
        return self.output_block(dpre[4], dpre[3], dpre[2], dpost[4], dpost[3], dpost[2])


class FusedEncUNet(_FusedEncoders):
    def __init__(self, args, n_class):
        super().__init__()
        self.compute_dtype = compute_dtype(args)
        enc_chn = self._build_encoders(args)
        self.no_skip = False
        dec_chn, self.dec_l1, self.dec_l2, self.dec_l3, self.dec_l4, self.dec_l5 = get_decoder(
            enc_chn, self.dilation, args.attention, args.dec_interp)
        self.output_block = OutputTemplate(n_class, args.deep_supervision, dec_chn, 1)

    def forward(self, data):
        _, posts = self._encode(self._prep(data))
        stages = [self.dec_l1, self.dec_l2, self.dec_l3, self.dec_l4, self.dec_l5]
        return self.output_block(*_decode(stages, posts, 1, False))


class ParallelUNet(_Net):
    def __init__(self, args, n_class):
        super().__init__()
        self.compute_dtype = compute_dtype(args)
        self.unet_pre = UNetTemplate(args)
        self.unet_post = UNetTemplate(args)
        self.output_block = OutputTemplate(n_class, args.deep_supervision, self.unet_pre.dec_chn, 2, args.interpolate,
                                           self.unet_pre.enc_chn[-1])

    def forward(self, data):
        data = self._prep(data)
        # reference quirk kept (unet.py:442-443): BOTH halves come from unet_pre on the PRE image
        first = self.unet_pre(data[:, :3])
        second = self.unet_pre(data[:, :3])
        return self.output_block(*first, *second)


class ParallelEncUNet(_Net, _TwinEncoderMixin):
    def __init__(self, args, n_class):
        super().__init__()
        self.compute_dtype = compute_dtype(args)
        self.use_ppm, self.use_aspp = args.ppm, args.aspp
        self.dilation, self.no_skip, self.interpolate = args.dilation, args.no_skip, args.interpolate
        self.enc_chn, self.enc_l1_pre, self.enc_l2_pre, self.enc_l3_pre, self.enc_l4_pre, self.enc_l5_pre = get_encoder(
            args.encoder, self.dilation)
        _, self.enc_l1_post, self.enc_l2_post, self.enc_l3_post, self.enc_l4_post, self.enc_l5_post = get_encoder(
            args.encoder, self.dilation)
        if self.use_ppm:
            self.ppm_pre = PPM(self.enc_chn[-1])
            self.ppm_post = PPM(self.enc_chn[-1])
        elif self.use_aspp:
            self.aspp_pre = ASPP(self.enc_chn[-1], self.dilation)
            self.aspp_post = ASPP(self.enc_chn[-1], self.dilation)
        self.dec_chn = None
        self.enc_chn = [2 * c for c in self.enc_chn]
        if not self.interpolate:
            self.dec_chn, self.dec_l1, self.dec_l2, self.dec_l3, self.dec_l4, self.dec_l5 = get_decoder(
                self.enc_chn, self.dilation, args.attention, self.no_skip, args.dec_interp)
        self.output_block = OutputTemplate(n_class, args.deep_supervision, self.dec_chn, 1, args.interpolate, self.enc_chn[-1])

    def forward_enc(self, data, pre):
        tag = "pre" if pre else "post"
        feats = []
        for i in range(1, 6):
            data = getattr(self, f"enc_l{i}_{tag}")(data)
            feats.append(data)
        return feats

    def forward(self, data):
        data = self._prep(data)
        pre, post = self.forward_enc(data[:, :3], True), self.forward_enc(data[:, 3:], False)
        if self.use_ppm:
            pre[4], post[4] = self.ppm_pre(pre[4]), self.ppm_post(post[4])
        elif self.use_aspp:
            pre[4], post[4] = self.aspp_pre(pre[4]), self.aspp_post(post[4])
        if self.interpolate:
            return self.output_block(pre[4], None, None, post[4])
        return self.output_block(*self._decode_cat(pre, post))


class DiffUNet(nn.Module):
    def __init__(self, args, n_class):
        super().__init__()
        self.unet = UNetLoc(args, in_channels=3, n_class=n_class)

    def forward(self, data):
        return self.unet(data[:, :3] - data[:, 3:])


class CatUNet(nn.Module):
    """The reference raises TypeError here (`"st" in encoder` on a module, unet.py:66); this build constructs the
    6-channel stem the code intended."""

    def __init__(self, args, n_class):
        super().__init__()
        self.unet = UNetLoc(args, in_channels=6, n_class=n_class)

    def forward(self, data):
        return self.unet(data)


def get_dmg_unet(args):
    table = {"siamese": SiameseUNet, "siameseEnc": SiameseEncUNet, "fused": FusedUNet, "fusedEnc": FusedEncUNet,
             "parallel": ParallelUNet, "parallelEnc": ParallelEncUNet, "diff": DiffUNet, "cat": CatUNet}
    return table[args.dmg_model](args, get_nclass(args))
