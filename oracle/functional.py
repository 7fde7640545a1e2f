"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Functional (stateless, dict-of-tensors) CPU restatement of the reference's segmentation hot path in plain
PyTorch fp32.  It travels to the GPU box (where /root/reference does not exist) and is the checker for the CUDA
path and the ``cpu_baseline`` / ``--impl reference`` leg of bench.py ("kind": "port").

Pinned (tests/test_oracle.py) against the reference's own modules executed verbatim in the build container
(oracle/refload.py) and against the committed tests/golden/*.pt those modules produced.

Every function cites the reference lines it follows.  ``P`` is a flat ``{state_dict_key: tensor}`` mapping using the
reference's checkpoint key names; BN running statistics in ``P`` are updated in place in training mode
exactly as nn.BatchNorm2d does.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------------------
# elementary blocks
# --------------------------------------------------------------------------------------------------------------


def _bn(P, k, x, training, momentum=0.1, eps=1e-5):
    """nn.BatchNorm2d semantics (train: batch stats + running update; eval: running stats)."""
    if training and (k + ".num_batches_tracked") in P:
        P[k + ".num_batches_tracked"] += 1
    return F.batch_norm(x, P[k + ".running_mean"], P[k + ".running_var"], P[k + ".weight"], P[k + ".bias"],
                        training, momentum, eps)


def conv_layer(P, k, x, training):
    """ConvLayer: 3x3 p1 no-bias conv -> BN -> LeakyReLU(0.01).  layers.py:89-100"""
    y = F.conv2d(x, P[k + ".conv.weight"], None, 1, 1)
    return F.leaky_relu(_bn(P, k + ".batch_norm", y, training), 0.01)


def conv_block(P, k, x, training):
    """ConvBlock = two ConvLayers.  layers.py:119-128"""
    return conv_layer(P, k + ".conv2", conv_layer(P, k + ".conv1", x, training), training)


def attention_layer(P, k, x, training):
    """AttentionLayer: 1x1 no-bias conv -> BN.  layers.py:68-77"""
    return _bn(P, k + ".batch_norm", F.conv2d(x, P[k + ".conv.weight"]), training)


def upsample_block(P, k, x, skip, training, attention, dec_interp=False):
    """UpsampleBlock.  layers.py:152-168 (transposed-conv flavour; --dec_interp: 3x3 conv + bias then bilinear x2, :153-154)"""
    if dec_interp:
        out = F.interpolate(F.conv2d(x, P[k + ".conv.weight"], P[k + ".conv.bias"], 1, 1), scale_factor=2, mode="bilinear",
                            align_corners=True)
    else:
        out = F.conv_transpose2d(x, P[k + ".conv_tranpose.conv.weight"], None, 2)  # layers.py:83,156 (attribute typo is the key)
    if skip is None:  # skip_channels == 0, layers.py:158-159
        return conv_block(P, k + ".conv_block", out, training)
    if attention:  # layers.py:161-166
        a = attention_layer(P, k + ".conv_o", out, training) + attention_layer(P, k + ".conv_s", skip, training)
        psi = attention_layer(P, k + ".psi", F.relu(a), training)
        skip = skip * torch.sigmoid(psi)
    return conv_block(P, k + ".conv_block", torch.cat((out, skip), 1), training)


def ppm(P, k, x, training):
    """PPM (layers.py:6-29): bins 1,2,3,6 -> 1x1 conv -> BN -> LeakyReLU -> bilinear (align_corners) back, cat, 1x1 conv + bias."""
    outs = [x]
    for i, b in enumerate((1, 2, 3, 6)):
        f = F.conv2d(F.adaptive_avg_pool2d(x, b), P[f"{k}.features.{i}.1.weight"])
        f = F.leaky_relu(_bn(P, f"{k}.features.{i}.2", f, training), 0.01)
        outs.append(F.interpolate(f, x.shape[2:], mode="bilinear", align_corners=True))
    return F.conv2d(torch.cat(outs, 1), P[k + ".conv.weight"], P[k + ".conv.bias"])


def aspp(P, k, x, training, dilation):
    """ASPP (layers.py:32-65): 1x1 and three dilated 3x3 (3, 6, 9 x dilation) conv -> BN -> LeakyReLU branches, cat."""
    outs = []
    for i, d in enumerate((1, 3 * dilation, 6 * dilation, 9 * dilation)):
        w = P[f"{k}.aspp{i + 1}.conv.weight"]
        y = F.conv2d(x, w) if i == 0 else F.conv2d(x, w, None, 1, d, d)
        outs.append(F.leaky_relu(_bn(P, f"{k}.aspp{i + 1}.bn", y, training), 0.01))
    return torch.cat(outs, 1)


# --------------------------------------------------------------------------------------------------------------
# encoders (external arithmetic restated)
# --------------------------------------------------------------------------------------------------------------

RESNEST = {"resnest50": ([3, 4, 6, 3], 32), "resnest101": ([3, 4, 23, 3], 64),
           "resnest200": ([3, 24, 36, 3], 64), "resnest269": ([3, 30, 48, 8], 64)}
RESNET = {"resnet50": [3, 4, 6, 3], "resnet101": [3, 4, 23, 3], "resnet152": [3, 8, 36, 3]}


def encoder_channels(encoder):
    """unet.py:49-54"""
    if "resnest" in encoder:
        return [64 if "50" in encoder else 128, 256, 512, 1024, 2048]
    return [64, 256, 512, 1024, 2048]


def _splat(P, k, x, training, dilation):
    """SplAtConv2d, radix 2, cardinality 1 (ResNeSt paper sec. 3; call site unet.py:52)."""
    y = F.conv2d(x, P[k + ".conv.weight"], None, 1, dilation, dilation, groups=2)
    y = F.relu(_bn(P, k + ".bn0", y, training))
    c = y.shape[1] // 2
    x0, x1 = y[:, :c], y[:, c:]
    gap = F.adaptive_avg_pool2d(x0 + x1, 1)
    gap = F.conv2d(gap, P[k + ".fc1.weight"], P[k + ".fc1.bias"])
    gap = F.relu(_bn(P, k + ".bn1", gap, training))
    att = F.conv2d(gap, P[k + ".fc2.weight"], P[k + ".fc2.bias"])
    b = att.shape[0]
    att = torch.softmax(att.view(b, 1, 2, c).transpose(1, 2), 1).reshape(b, 2 * c, 1, 1)
    return att[:, :c] * x0 + att[:, c:] * x1


def _resnest_block(P, k, x, training, stride, dilation, is_first, has_down):
    out = F.relu(_bn(P, k + ".bn1", F.conv2d(x, P[k + ".conv1.weight"]), training))
    out = _splat(P, k + ".conv2", out, training, dilation)
    if stride > 1 or is_first:  # avd, after the SplAt conv (avd_first=False)
        out = F.avg_pool2d(out, 3, stride, 1)
    out = _bn(P, k + ".bn3", F.conv2d(out, P[k + ".conv3.weight"]), training)
    res = x
    if has_down:  # avg_down shortcut: AvgPool(stride, ceil, no pad count) -> 1x1 -> BN
        kk = stride if dilation == 1 else 1
        if kk > 1:
            res = F.avg_pool2d(res, kk, kk, 0, ceil_mode=True, count_include_pad=False)
        res = _bn(P, k + ".downsample.2", F.conv2d(res, P[k + ".downsample.1.weight"]), training)
    return F.relu(out + res)


def _resnet_block(P, k, x, training, stride, dilation, has_down):
    """torchvision Bottleneck (stride on the 3x3, v1.5).  call site unet.py:54-63"""
    out = F.relu(_bn(P, k + ".bn1", F.conv2d(x, P[k + ".conv1.weight"]), training))
    out = F.relu(_bn(P, k + ".bn2", F.conv2d(out, P[k + ".conv2.weight"], None, stride, dilation, dilation), training))
    out = _bn(P, k + ".bn3", F.conv2d(out, P[k + ".conv3.weight"]), training)
    res = x
    if has_down:
        res = _bn(P, k + ".downsample.1", F.conv2d(x, P[k + ".downsample.0.weight"], None, stride), training)
    return F.relu(out + res)


def _stage_plan(dilation):
    """(stride, first-block dilation, other-block dilation) of layer1..4 for --dilation 1|2|4."""
    if dilation == 1:
        return [(1, 1, 1), (2, 1, 1), (2, 1, 1), (2, 1, 1)]
    if dilation == 2:
        return [(1, 1, 1), (2, 1, 1), (2, 1, 1), (1, 1, 2)]
    return [(1, 1, 1), (2, 1, 1), (1, 1, 2), (1, 2, 4)]


def encoder_forward(P, k, x, training, encoder, dilation=1):
    """get_encoder's five stages (unet.py:80-84) applied in order; returns (enc1..enc5).

    ``k`` is the prefix in front of ``enc_l1`` .. ``enc_l5`` (e.g. ``"unet."``); ``suffix`` variants are handled by callers.
    """
    return encoder_forward_named(P, [k + f"enc_l{i}" for i in range(1, 6)], x, training, encoder, dilation)


def encoder_forward_named(P, names, x, training, encoder, dilation=1):
    feats = []
    for i in range(5):
        x = encoder_stage(P, names[i], i, x, training, encoder, dilation)
        feats.append(x)
    return feats


def encoder_stage(P, name, i, x, training, encoder, dilation=1):
    """Stage ``i`` (0-based) of get_encoder: 0 = stem+bn1+relu, 1 = maxpool+layer1, 2..4 = layer2..4."""
    is_st = "resnest" in encoder
    blocks = RESNEST[encoder][0] if is_st else RESNET[encoder]
    if i == 0:
        if is_st:  # deep stem Sequential idx 0,1,3,4,6 then bn1 (name.1)
            s = name + ".0"
            x = F.relu(_bn(P, s + ".1", F.conv2d(x, P[s + ".0.weight"], None, 2, 1), training))
            x = F.relu(_bn(P, s + ".4", F.conv2d(x, P[s + ".3.weight"], None, 1, 1), training))
            x = F.conv2d(x, P[s + ".6.weight"], None, 1, 1)
        else:
            x = F.conv2d(x, P[name + ".0.weight"], P.get(name + ".0.bias"), 2, 3)
        return F.relu(_bn(P, name + ".1", x, training))
    plan = _stage_plan(dilation)
    stride, d_first, d_rest = plan[i - 1]
    if is_st:
        # ResNeSt's dilation==2 path dilates layer4 blocks by 2 except the first (upstream _make_layer)
        pre = name + ".1." if i == 1 else name + "."
        if i == 1:
            x = F.max_pool2d(x, 3, 2, 1)
        for b in range(blocks[i - 1]):
            if b == 0:
                x = _resnest_block(P, pre + "0", x, training, stride, d_first, is_first=(i != 1), has_down=True)
            else:
                x = _resnest_block(P, pre + str(b), x, training, 1, d_rest, False, False)
        return x
    # torchvision replace_stride_with_dilation: first block keeps the previous dilation, the rest use the new one
    # -- the same (stride, first, rest) plan as ResNeSt's.
    pre = name + ".1." if i == 1 else name + "."
    if i == 1:
        x = F.max_pool2d(x, 3, 2, 1)
    for b in range(blocks[i - 1]):
        if b == 0:
            x = _resnet_block(P, pre + "0", x, training, stride, d_first, has_down=True)
        else:
            x = _resnet_block(P, pre + str(b), x, training, 1, d_rest, False)
    return x


# --------------------------------------------------------------------------------------------------------------
# U-Net assembly
# --------------------------------------------------------------------------------------------------------------


def decoder_forward(P, names, encs, training, attention, dilation=1, no_skip=False, dec_interp=False):
    """Decoder half of UNetTemplate.forward (unet.py:153-170); names = dec_l1..dec_l5 prefixes.  --dilation 2 / 4 drop the
    first one / two stages, --no_skip feeds no encoder features."""
    enc1, enc2, enc3, enc4, enc5 = encs
    first = {1: 0, 2: 1, 4: 2}[dilation]
    skips = [enc4, enc3, enc2, enc1, None]
    x, outs = enc5, {}
    for i in range(first, 5):
        x = upsample_block(P, names[i], x, None if no_skip else skips[i], training, attention, dec_interp)
        outs[i] = x
    return outs[4], outs[3], outs[2]


def _opt(args, name, default=False):
    return getattr(args, name, default)


def _context(P, k, enc5, training, args, suffix=""):
    """--ppm / --aspp on the last encoder stage (unet.py:144-147)."""
    if _opt(args, "ppm"):
        return ppm(P, k + "ppm" + suffix, enc5, training)
    if _opt(args, "aspp"):
        return aspp(P, k + "aspp" + suffix, enc5, training, _opt(args, "dilation", 1))
    return enc5


def unet_template(P, k, x, training, args):
    """UNetTemplate.forward (unet.py:136-172) -> (dec5, dec4, dec3); --interpolate returns (enc5, None, None)."""
    dil = _opt(args, "dilation", 1)
    encs = encoder_forward(P, k, x, training, args.encoder, dil)
    encs[4] = _context(P, k, encs[4], training, args)
    if _opt(args, "interpolate"):
        return encs[4], None, None
    return decoder_forward(P, [k + f"dec_l{i}" for i in range(1, 6)], encs, training, args.attention, dil,
                           _opt(args, "no_skip"), _opt(args, "dec_interp"))


def output_block(P, k, x, training, interpolate=False):
    """OutputBlock.forward (layers.py:182-189): 1x1 conv (+ bias | + the CORAL rank biases), optional bilinear resize."""
    if (k + ".bias") in P:  # coral head: 1-channel conv without bias + a (3, 1, 1) bias parameter (layers.py:175-178)
        out = F.conv2d(x, P[k + ".conv.weight"]) + P[k + ".bias"]
    else:
        out = F.conv2d(x, P[k + ".conv.weight"], P[k + ".conv.bias"])
    if interpolate:
        out = F.interpolate(out, (512, 512) if training else (1024, 1024), mode="bilinear", align_corners=True)
    return out


def output_template(P, k, dec5, dec4, dec3, training, deep_supervision, interpolate=False):
    """OutputTemplate.forward (unet.py:191-197): heads; DS heads only in training (and never with --interpolate, :181-182)."""
    out = output_block(P, f"{k}.output_block", dec5, training, interpolate)
    if training and deep_supervision and not interpolate:
        return [out, output_block(P, f"{k}.output_block_ds4", dec4, training), output_block(P, f"{k}.output_block_ds3", dec3, training)]
    return out


def _cat(a, b):
    """concat (unet.py:17-18)."""
    return None if a is None or b is None else torch.cat([a, b], 1)


def unet_loc(P, x, training, args, prefix=""):
    """UNetLoc.forward (unet.py:212-215)."""
    d5, d4, d3 = unet_template(P, prefix + "unet.", x, training, args)
    return output_template(P, prefix + "output_block", d5, d4, d3, training, args.deep_supervision, _opt(args, "interpolate"))


def siamese_unet(P, x, training, args, prefix=""):
    """SiameseUNet.forward (unet.py:231-236): the SAME U-Net on pre then post (separate BN statistics), cat, heads."""
    pre = unet_template(P, prefix + "unet.", x[:, :3], training, args)
    post = unet_template(P, prefix + "unet.", x[:, 3:], training, args)
    d5, d4, d3 = (_cat(a, b) for a, b in zip(pre, post))
    return output_template(P, prefix + "output_block", d5, d4, d3, training, args.deep_supervision, _opt(args, "interpolate"))


def _twin_decode(P, p, pre, post, training, args):
    encs = [_cat(a, b) for a, b in zip(pre, post)]
    d5, d4, d3 = decoder_forward(P, [p + f"dec_l{i}" for i in range(1, 6)], encs, training, args.attention,
                                 _opt(args, "dilation", 1), _opt(args, "no_skip"), _opt(args, "dec_interp"))
    return output_template(P, p + "output_block", d5, d4, d3, training, args.deep_supervision)


def siamese_enc_unet(P, x, training, args, prefix=""):
    """SiameseEncUNet.forward (unet.py:275-314): one shared ENCODER on pre and post, features concatenated, one decoder."""
    p, dil = prefix, _opt(args, "dilation", 1)
    feats = []
    for part in (x[:, :3], x[:, 3:]):
        encs = encoder_forward(P, p, part, training, args.encoder, dil)
        encs[4] = _context(P, p, encs[4], training, args)
        feats.append(encs)
    return _twin_decode(P, p, feats[0], feats[1], training, args)


def parallel_unet(P, x, training, args, prefix=""):
    """ParallelUNet.forward (unet.py:441-446) with its quirk: BOTH halves are unet_pre on the PRE image."""
    first = unet_template(P, prefix + "unet_pre.", x[:, :3], training, args)
    second = unet_template(P, prefix + "unet_pre.", x[:, :3], training, args)
    d5, d4, d3 = (_cat(a, b) for a, b in zip(first, second))
    return output_template(P, prefix + "output_block", d5, d4, d3, training, args.deep_supervision, _opt(args, "interpolate"))


def parallel_enc_unet(P, x, training, args, prefix=""):
    """ParallelEncUNet.forward (unet.py:497-540): separate pre / post encoders, features concatenated, one decoder."""
    p, dil = prefix, _opt(args, "dilation", 1)
    feats = []
    for part, tag in ((x[:, :3], "pre"), (x[:, 3:], "post")):
        encs = encoder_forward_named(P, [f"{p}enc_l{i}_{tag}" for i in range(1, 6)], part, training, args.encoder, dil)
        encs[4] = _context(P, p, encs[4], training, args, "_" + tag)
        feats.append(encs)
    if _opt(args, "interpolate"):
        return output_template(P, p + "output_block", _cat(feats[0][4], feats[1][4]), None, None, training, False, True)
    return _twin_decode(P, p, feats[0], feats[1], training, args)


def diff_unet(P, x, training, args, prefix=""):
    """DiffUNet.forward (unet.py:548-551)."""
    return unet_loc(P, x[:, :3] - x[:, 3:], training, args, prefix + "unet.")


def cat_unet(P, x, training, args, prefix=""):
    """CatUNet.forward (unet.py:558-560) with the 6-channel stem the reference intended (its constructor raises)."""
    return unet_loc(P, x, training, args, prefix + "unet.")


def _fusion(P, k, pre, post, training):
    """FusionBlock tail (layers.py:114-116): cat -> two ConvLayer(2C->C)."""
    f = torch.cat([pre, post], 1)
    return conv_layer(P, k + ".conv_pre", f, training), conv_layer(P, k + ".conv_post", f, training)


def fused_unet(P, x, training, args, prefix=""):
    """FusedUNet.forward (unet.py:360-376): twin encoders/decoders with cross-fusion after each of 10 stages."""
    p = prefix
    pre, post = x[:, :3], x[:, 3:]
    e_pre, e_post = [], []
    for i in range(5):
        pre = encoder_stage(P, f"{p}enc_l{i+1}_pre", i, pre, training, args.encoder, 1)
        post = encoder_stage(P, f"{p}enc_l{i+1}_post", i, post, training, args.encoder, 1)
        pre, post = _fusion(P, f"{p}fusion_block{i+1}", pre, post, training)
        e_pre.append(pre)
        e_post.append(post)
    d_pre, d_post = [], []
    for i in range(5):
        sk_pre = e_pre[3 - i] if i < 4 else None
        sk_post = e_post[3 - i] if i < 4 else None
        pre = upsample_block(P, f"{p}dec_l{i+1}_pre", pre, sk_pre, training, args.attention)
        post = upsample_block(P, f"{p}dec_l{i+1}_post", post, sk_post, training, args.attention)
        pre, post = _fusion(P, f"{p}fusion_block_dec{i+1}", pre, post, training)
        d_pre.append(pre)
        d_post.append(post)
    d5, d4, d3 = (torch.cat([d_pre[j], d_post[j]], 1) for j in (4, 3, 2))
    return output_template(P, p + "output_block", d5, d4, d3, training, args.deep_supervision)


def fused_enc_unet(P, x, training, args, prefix=""):
    """FusedEncUNet.forward (unet.py:411-426): fused twin encoders, ONE decoder fed with the post branch."""
    p = prefix
    pre, post = x[:, :3], x[:, 3:]
    e_post = []
    for i in range(5):
        pre = encoder_stage(P, f"{p}enc_l{i+1}_pre", i, pre, training, args.encoder, 1)
        post = encoder_stage(P, f"{p}enc_l{i+1}_post", i, post, training, args.encoder, 1)
        pre, post = _fusion(P, f"{p}fusion_block{i+1}", pre, post, training)
        e_post.append(post)
    d5, d4, d3 = decoder_forward(P, [p + f"dec_l{i}" for i in range(1, 6)], e_post, training, args.attention, 1,
                                 _opt(args, "dec_interp"))  # dec_interp sits in get_decoder's no_skip slot (unet.py:398-400)
    return output_template(P, p + "output_block", d5, d4, d3, training, args.deep_supervision)


def model_forward(P, x, training, args, prefix=""):
    """Model.__init__ dispatch (plt.py:26; get_dmg_unet unet.py:29-42)."""
    if args.type == "pre":
        return unet_loc(P, x, training, args, prefix)
    table = {"siamese": siamese_unet, "siameseEnc": siamese_enc_unet, "fused": fused_unet, "fusedEnc": fused_enc_unet,
             "parallel": parallel_unet, "parallelEnc": parallel_enc_unet, "diff": diff_unet, "cat": cat_unet}
    if args.dmg_model not in table:
        raise NotImplementedError(args.dmg_model)
    return table[args.dmg_model](P, x, training, args, prefix)


def tta_forward(P, x, args, prefix=""):
    """Model.forward with --tta (plt.py:42-48): mean of logits over {id, flip H, flip W, flip HW}."""
    pred = model_forward(P, x, False, args, prefix)
    for dims in ([2], [3], [2, 3]):
        pred = pred + torch.flip(model_forward(P, torch.flip(x, dims), False, args, prefix), dims)
    return pred / 4


# --------------------------------------------------------------------------------------------------------------
# losses (loss.py) -- MONAI 0.4.0 formulas restated
# --------------------------------------------------------------------------------------------------------------


def dice_loss(logits2d, target1d, include_background):
    """MonaiLoss('dice') on flattened (M, C) logits / (M,) labels: batch=True reduces over everything.  loss.py:12-13,17-20"""
    p = torch.softmax(logits2d, 1)
    t = F.one_hot(target1d, logits2d.shape[1]).to(p.dtype)
    if not include_background:
        p, t = p[:, 1:], t[:, 1:]
    inter, denom = (p * t).sum(0), p.sum(0) + t.sum(0)
    return (1 - (2 * inter + 1e-5) / (denom + 1e-5)).mean()


def focal_loss(logits2d, target1d, gamma=2.0):
    """MonaiLoss('focal'), gamma 2, mean over pixels.  loss.py:11,21"""
    logpt = F.log_softmax(logits2d, 1).gather(1, target1d[:, None])[:, 0]
    return (-(1 - logpt.exp()) ** gamma * logpt).mean()


def ce_loss(logits2d, target1d):
    return F.cross_entropy(logits2d, target1d)


CORAL_LEVELS = torch.tensor([[0, 0, 0], [1, 0, 0], [1, 1, 0], [1, 1, 1]], dtype=torch.float32)


def coral_loss(logits2d, target1d):
    """CORAL (loss.py:54-65) on flattened (M, 3) rank logits / (M,) labels 0..3."""
    levels = CORAL_LEVELS.to(logits2d.device)[target1d]
    logpt = F.logsigmoid(logits2d)
    return -torch.mean(torch.sum(logpt * levels + (logpt - logits2d) * (1 - levels), dim=1))


def convert_to_labels(loss_str, logits):
    """utils/f1.py:7-15."""
    if loss_str == "mse":
        return torch.round(F.relu(logits[:, 0])).clamp_max(3) + 1
    if loss_str == "coral":
        return torch.sum(torch.sigmoid(logits) > 0.5, dim=1) + 1
    return torch.argmax(logits, dim=1) + 1


def loss_forward(y_pred, y_true, loss_str, post):
    """Loss.forward (loss.py:85-101).  `post` keeps only pixels with label > 0 and shifts labels by -1.

    'ohem' is the reference's effective behaviour: loss.py:45 slices the (values, indices) tuple so no negative is
    ever dropped and the result is sum(CE)/N == mean CE.
    """
    c = y_pred.shape[1]
    flat = y_pred.permute(0, 2, 3, 1).reshape(-1, c)
    tgt = y_true.reshape(-1).long()
    if post:
        keep = tgt > 0
        flat, tgt = flat[keep], tgt[keep] - 1
    if loss_str == "mse":  # loss.py:92-94: relu on channel 0, float targets, nn.MSELoss
        return F.mse_loss(F.relu(flat[:, 0]), tgt.float())
    if loss_str == "coral":
        return coral_loss(flat, tgt)
    total = 0
    for name in loss_str.split("+"):
        if name == "dice":
            total = total + dice_loss(flat, tgt, include_background=(c != 2))
        elif name == "focal":
            total = total + focal_loss(flat, tgt)
        elif name in ("ce", "ohem"):
            total = total + ce_loss(flat, tgt)
        else:
            raise NotImplementedError(name)
    return total


def compute_loss(preds, label, loss_str, post, deep_supervision):
    """Model.compute_loss (plt.py:69-77): 1, 1/2, 1/4 weights on nearest-downsampled labels, normalised."""
    if not deep_supervision:
        return loss_forward(preds, label, loss_str, post)
    loss = loss_forward(preds[0], label, loss_str, post)
    for i, pred in enumerate(preds[1:]):
        f = label.shape[-1] // pred.shape[-1]
        loss = loss + 0.5 ** (i + 1) * loss_forward(pred, label[:, ::f, ::f], loss_str, post)  # F.interpolate nearest on u8
    return loss / (2 - 2 ** (-len(preds)))


# --------------------------------------------------------------------------------------------------------------
# metric, post-process, loader normalise
# --------------------------------------------------------------------------------------------------------------


def f1_counters(logits, targets, n_class):
    """F1.update (utils/f1.py:28-42): returns (tp, fp, fn) int64 arrays of length n_class-1."""
    pred = torch.argmax(torch.softmax(logits, 1), 1)
    tgt = targets.long()
    if n_class == 5:
        pred = pred + 1
        keep = tgt > 0
        pred, tgt = pred[keep], tgt[keep]
    tp, fp, fn = [], [], []
    for c in range(1, n_class):
        tp.append(int(((pred == c) & (tgt == c)).sum()))
        fn.append(int(((pred != c) & (tgt == c)).sum()))
        fp.append(int(((pred == c) & (tgt != c)).sum()))
    return np.array(tp), np.array(fp), np.array(fn)


def f1_compute(tp, fp, fn):
    """F1.compute (utils/f1.py:44-49)."""
    tp, fp, fn = (np.asarray(a, np.float32) for a in (tp, fp, fn))
    f1 = 200 * tp / (2 * tp + fp + fn)
    if len(tp) == 4:
        return np.float32(4 / sum((f + np.float32(1e-6)) ** -1 for f in f1)), f1
    return f1, None


def post_process(loc, dmg):
    """utils/post_process.py:27-38 without --components/--dilate: (pre u8, post u8) label maps."""
    post = np.argmax(dmg, axis=0) + 1 if dmg.shape[0] == 4 else dmg
    idx = np.logical_or(loc > 0.3, np.logical_and(loc > 0.1, post > 1))
    pre = np.zeros(loc.shape)
    pre[idx] = 1
    post = post * pre
    return pre.astype(np.uint8), post.astype(np.uint8)


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def normalize_tile(img_u8_hwc):
    """A.Normalize() + HWC->CHW as TestDataset does (pytorch_loader.py:63,165,170): (x - 255*mean) * (1/(255*std)) in f32.

    cv2 hands over BGR but the RGB ImageNet constants are applied by channel position (reference quirk).
    """
    mean = np.array(IMAGENET_MEAN, np.float32) * 255.0
    inv = np.reciprocal(np.array(IMAGENET_STD, np.float32) * 255.0, dtype=np.float32)
    x = (img_u8_hwc.astype(np.float32) - mean) * inv
    return np.transpose(x, (2, 0, 1))


def noam_lr(step, warmup_steps, total_steps, init_lr, max_lr, final_lr):
    """NoamLR.step (utils/scheduler.py:45-59) for one param group."""
    if step <= warmup_steps:
        return init_lr + step * (max_lr - init_lr) / warmup_steps
    if step <= total_steps:
        gamma = (final_lr / max_lr) ** (1 / (total_steps - warmup_steps))
        return max_lr * gamma ** (step - warmup_steps)
    return final_lr


def canonical_key(key):
    """FusedUNet registers each stage twice (unet.py:326,333: enc_l1_pre IS fusion_block1.pre_conv): map the alias
    spelling of a state_dict key to the canonical one."""
    import re
    key = re.sub(r"fusion_block_dec(\d)\.(pre|post)_conv\.", r"dec_l\1_\2.", key)
    return re.sub(r"fusion_block(\d)\.(pre|post)_conv\.", r"enc_l\1_\2.", key)


def deterministic_state(shapes, seed=1):
    """Order-independent deterministic fill used by the golden fixtures: every tensor is drawn from its own
    generator seeded by crc32(key)^seed, so the reference modules, this oracle and the CUDA path can be given
    identical weights without shipping a 170 MB state_dict.  ``shapes``: {key: (shape, dtype)}.
    """
    import zlib
    out = {}
    for key in sorted(shapes):
        shape, dtype = shapes[key]
        canon = canonical_key(key)  # both aliases of a shared module receive the same values
        g = torch.Generator().manual_seed((zlib.crc32(canon.encode()) ^ seed) & 0x7FFFFFFF)
        if key.endswith("num_batches_tracked"):
            t = torch.zeros(shape, dtype=dtype)
        elif key.endswith("running_mean"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("running_var"):
            t = 0.5 + torch.rand(shape, generator=g)
        elif len(shape) == 1 and key.endswith("weight"):  # BN gamma
            t = 0.5 + torch.rand(shape, generator=g)
        elif key.endswith("bias"):
            t = 0.1 * torch.randn(shape, generator=g)
        else:  # conv / convT weight: He-style fan-in scaling keeps activations O(1) through 100+ layers
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        if ".fc2." in key:  # SplAt attention logits: moderate like a trained net's, not a saturated hard switch
            t = t * 0.25
        out[key] = t.to(dtype)
    return out


# --------------------------------------------------------------------------------------------------------------
# train-time augmentation (pytorch_loader.py:57-63,73-92,109-115,124-148), restated per OUTPUT pixel in float32 exactly as the
# device kernel evaluates it (xview2_b200/csrc/augment.cu); cv2 / albumentations semantics are cited inline
# --------------------------------------------------------------------------------------------------------------
def _hash32(x):
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7feb352d)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846ca68b)
    x ^= x >> np.uint32(16)
    return x


def counter_normal(seed, idx):
    """N(0, 1) from a counter: Box-Muller on two hashed 24-bit uniforms (float32)."""
    idx = idx.astype(np.uint32)
    seed = np.uint32(seed)
    with np.errstate(over="ignore"):
        a = _hash32(idx * np.uint32(2) + np.uint32(0x9e3779b9) * seed)
        b = _hash32(idx * np.uint32(2) + np.uint32(1) + np.uint32(0x85ebca6b) * seed)
    u1 = ((a >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 16777216.0)
    u2 = (b >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return np.sqrt(np.float32(-2.0) * np.log(u1)) * np.cos(np.float32(6.28318530717958647692) * u2)


def _cubic_weights(t):
    """Keys cubic, A = -0.75 (cv2 INTER_CUBIC), float32, same operation order as the device kernel."""
    A = np.float32(-0.75)
    t = t.astype(np.float32)
    t1, u = t + np.float32(1), np.float32(1) - t
    w0 = ((A * t1 + np.float32(3.75)) * t1 + np.float32(-6.0)) * t1 + np.float32(3.0)
    w1 = ((np.float32(1.25) * t + np.float32(-2.25)) * t) * t + np.float32(1)
    w2 = ((np.float32(1.25) * u + np.float32(-2.25)) * u) * u + np.float32(1)
    w3 = ((np.float32(1) - w0) - w1) - w2
    return [w0, w1, w2, w3]


def augment_restatement(pre, post, mask, P, origin, crop=512):
    """One sample.  pre / post: uint8 (H, W, 3) (post may be None), mask uint8 (H, W), P: the 16 host-drawn floats of
    include/xv2.h, origin (x0, y0) in the scaled image.  Returns (uint8 (crop, crop, 3|6) augmented bytes, float32 normalised
    (crop, crop, 3|6), uint8 (crop, crop) mask).

    RandomScale: cv2.resize(dsize) -> scale = src / dst, cubic fx = (dx + .5) scale - .5 with replicated borders, nearest
    sx = min(floor(dx scale), src - 1) (pytorch_loader.py:58,110); CropNonEmptyMaskIfExists / flips are index arithmetic
    (:57,59-60); GaussNoise: clip(img + N(0, var)) truncated to uint8 per image (:61, intensity_aug :45-51);
    RandomBrightnessContrast: LUT clip(x alpha + beta 255) (:62); Normalize (:63)."""
    P = np.asarray(P, np.float32)
    sh, sw = mask.shape
    x0, y0 = int(origin[0]), int(origin[1])
    ys, xs = np.meshgrid(np.arange(crop), np.arange(crop), indexing="ij")
    xc = (crop - 1 - xs if P[6] != 0 else xs) + x0
    yc = (crop - 1 - ys if P[7] != 0 else ys) + y0
    zoom = P[15] != 0
    ifx, ify = sw / float(int(P[2])), sh / float(int(P[3]))  # cv2: double scale from the integer sizes
    if zoom:
        mx = np.minimum(np.floor(xc.astype(np.float64) * ifx).astype(np.int64), sw - 1)
        my = np.minimum(np.floor(yc.astype(np.float64) * ify).astype(np.int64), sh - 1)
    else:
        mx, my = xc, yc
    mask_out = mask[my, mx]
    imgs = [pre] + ([post] if post is not None else [])
    out_u8 = np.zeros((crop, crop, 3 * len(imgs)), np.uint8)
    if zoom:
        fx = ((xc.astype(np.float64) + 0.5) * ifx - 0.5).astype(np.float32)
        fy = ((yc.astype(np.float64) + 0.5) * ify - 0.5).astype(np.float32)
        sx, sy = np.floor(fx).astype(np.int64), np.floor(fy).astype(np.int64)
        wx, wy = _cubic_weights(fx - sx.astype(np.float32)), _cubic_weights(fy - sy.astype(np.float32))
        ix = [np.clip(sx - 1 + k, 0, sw - 1) for k in range(4)]
        iy = [np.clip(sy - 1 + k, 0, sh - 1) for k in range(4)]
    pix = (ys * crop + xs).astype(np.int64)  # output pixel index within the sample (the caller adds the sample offset)
    for im, img in enumerate(imgs):
        sigma, alpha, beta = P[8 + im], P[10 + 2 * im], P[11 + 2 * im]
        for ch in range(3):
            if zoom:
                acc = np.zeros((crop, crop), np.float32)
                for r in range(4):
                    row = np.zeros((crop, crop), np.float32)
                    for k in range(4):
                        row = row + wx[k] * img[iy[r], ix[k], ch].astype(np.float32)
                    acc = acc + wy[r] * row
                v = np.clip(np.rint(acc), 0, 255).astype(np.float32)
            else:
                v = img[yc, xc, ch].astype(np.float32)
            if sigma > 0:
                idx = ((pix + augment_restatement.sample_base) * 2 + im) * 3 + ch
                v = np.floor(np.clip(v + sigma * counter_normal(int(P[14]), idx), 0, 255)).astype(np.float32)
            if alpha != 1 or beta != 0:
                v = np.floor(np.clip(v * alpha + beta * np.float32(255), 0, 255)).astype(np.float32)
            out_u8[:, :, im * 3 + ch] = v.astype(np.uint8)
    mean = np.array(IMAGENET_MEAN * len(imgs), np.float32) * np.float32(255)
    inv = np.float32(1) / (np.array(IMAGENET_STD * len(imgs), np.float32) * np.float32(255))
    return out_u8, (out_u8.astype(np.float32) - mean) * inv, mask_out


augment_restatement.sample_base = 0  # index of the sample's first output pixel in the batch (i = img * crop^2 + pix on the device)


def crop_origin_restatement(mask, P, u, crop=512):
    """CropNonEmptyMaskIfExists.update_params (albumentations 0.5.1) on the nearest-scaled mask with the three uniforms the
    device kernel consumes: k = floor(u0 count)-th non-zero pixel (row-major), minus floor(u1 cw) / floor(u2 ch), clipped."""
    P = np.asarray(P, np.float32)
    sh, sw = mask.shape
    ws, hs = int(P[2]), int(P[3])
    if P[15] != 0:
        mx = np.minimum(np.floor(np.arange(ws, dtype=np.float64) * (sw / float(ws))).astype(np.int64), sw - 1)
        my = np.minimum(np.floor(np.arange(hs, dtype=np.float64) * (sh / float(hs))).astype(np.int64), sh - 1)
        scaled = mask[my][:, mx]
    else:
        scaled = mask
    yy, xx = np.nonzero(scaled)
    u = np.asarray(u, np.float32)
    if len(yy):
        k = min(int(np.floor(u[0] * np.float32(len(yy)))), len(yy) - 1)
        px = int(np.clip(xx[k] - int(np.floor(u[1] * np.float32(crop))), 0, ws - crop))
        py = int(np.clip(yy[k] - int(np.floor(u[2] * np.float32(crop))), 0, hs - crop))
    else:
        px = min(int(np.floor(u[1] * np.float32(ws - crop + 1))), ws - crop)
        py = min(int(np.floor(u[2] * np.float32(hs - crop + 1))), hs - crop)
    return px, py
