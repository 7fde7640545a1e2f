"""Debug aid (GPU box): walks the ResNeSt encoder block by block / op by op in eval mode, fp32, against torch on CPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from oracle import functional as OF
from tests.helpers import golden_inputs, golden_state, load_golden, rel_err
from xview2_b200 import ops
from xview2_b200.lib import ACT_NONE, ACT_RELU
from xview2_b200.model.layers import run_conv
from xview2_b200.model.encoders import _sub

torch.backends.cudnn.allow_tf32 = False
fx = load_golden("c2_resnest50_pre")
ns = fx["ns"]
ns.precision = 32
from xview2_b200.model.unet import UNetLoc

model = UNetLoc(ns)
state = golden_state(fx)
model.load_state_dict(state, strict=True)
model = model.cuda().eval()
x, y = golden_inputs(fx)
P = {k: v.clone() for k, v in state.items()}
with torch.no_grad():
    encs = OF.encoder_forward(P, "unet.", x, False, ns.encoder, 1)
    xin = encs[2]  # input of enc_l4
    cur_ref = xin
    cur = ops.nhwc(xin.cuda())
    for b, blk in enumerate(model.unet.enc_l4._modules.values()):
        k = f"unet.enc_l4.{b}"
        stride = 2 if b == 0 else 1
        # ours, op by op
        o1 = ops.batch_norm_act(run_conv(blk.conv1, cur), blk.bn1, ACT_RELU)
        r1 = F.relu(OF._bn(P, k + ".bn1", F.conv2d(cur_ref, P[k + ".conv1.weight"]), False))
        print(b, "conv1+bn1", rel_err(o1, r1))
        sp = blk.conv2
        o2 = ops.batch_norm_act(run_conv(sp.conv, o1), sp.bn0, ACT_RELU)
        r2 = F.relu(OF._bn(P, k + ".conv2.bn0", F.conv2d(r1, P[k + ".conv2.conv.weight"], None, 1, 1, 1, groups=2), False))
        print(b, "radix conv+bn0", rel_err(o2, r2), "(same input:", rel_err(ops.batch_norm_act(run_conv(sp.conv, ops.nhwc(r1.cuda())), sp.bn0, ACT_RELU), r2), ")")
        o3 = ops.split_attention(o2, sp.fc1, sp.bn1, sp.fc2)
        r3 = OF._splat(P, k + ".conv2", r1, False, 1)
        print(b, "split attention", rel_err(o3, r3), "(same input:", rel_err(ops.split_attention(ops.nhwc(r2.cuda()), sp.fc1, sp.bn1, sp.fc2), r3), ")")
        if blk.avd_stride:
            o3 = ops.avg_pool2d(o3, 3, blk.avd_stride, 1)
            r3 = F.avg_pool2d(r3, 3, stride, 1)
            print(b, "avd", rel_err(o3, r3))
        res, rres = cur, cur_ref
        if blk.downsample is not None:
            if blk.down_pool > 1:
                res = ops.avg_pool2d(res, blk.down_pool, blk.down_pool, 0, ceil_mode=True, count_include_pad=False)
                rres = F.avg_pool2d(rres, stride, stride, 0, ceil_mode=True, count_include_pad=False)
                print(b, "down pool", rel_err(res, rres))
            res = ops.batch_norm_act(run_conv(_sub(blk.downsample, "1"), res), _sub(blk.downsample, "2"), ACT_NONE)
            rres = OF._bn(P, k + ".downsample.2", F.conv2d(rres, P[k + ".downsample.1.weight"]), False)
            print(b, "down conv+bn", rel_err(res, rres))
        o4 = ops.batch_norm_act(run_conv(blk.conv3, o3), blk.bn3, ACT_RELU, residual=res)
        r4 = F.relu(OF._bn(P, k + ".bn3", F.conv2d(r3, P[k + ".conv3.weight"]), False) + rres)
        print(b, "conv3+bn3+res", rel_err(o4, r4), "whole block (module):", rel_err(blk(cur), r4))
        cur, cur_ref = o4, r4
