"""One launch of every kernel family of the training step at its BASELINE-config-2 shape (batch 8, 1024^2 tiles), bracketed by
cudaProfilerStart/Stop, so that ONE short `ncu --set full --profile-from-start off` run yields a tensor-pipe / DRAM row per kernel
 without replaying the 526-launch step.   python tools/kernel_zoo.py [--list]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn as nn

from xview2_b200 import lib, ops
from xview2_b200.lib import ACT_LRELU, ACT_RELU

ap = argparse.ArgumentParser()
ap.add_argument("--list", action="store_true")
ap.add_argument("--only", default="", help="comma-separated substrings: run only the cases whose name contains one of them")
a = ap.parse_args()
torch.cuda.set_device(0)
lib.init(0)
CL = torch.channels_last
g = torch.Generator(device="cuda").manual_seed(1)


def act(n, c, h, w, grad=True):
    t = torch.randn(n, c, h, w, device="cuda", generator=g).to(torch.bfloat16).contiguous(memory_format=CL)
    return t.requires_grad_(grad)


def conv(cin, cout, k, groups=1):
    m = nn.Conv2d(cin, cout, k, padding=k // 2, groups=groups, bias=False).cuda()
    m.weight.data = m.weight.data.contiguous(memory_format=CL)
    return m


def bn(c):
    return nn.BatchNorm2d(c).cuda().train()


cases = []  # (name, thunk)


def conv_case(name, n, cin, c2, h, w, cout, k, groups=1):
    x, x2 = act(n, cin, h, w), (act(n, c2, h, w) if c2 else None)
    m = conv(cin + c2, cout, k, groups)

    def run():
        y, _ = ops.conv2d_stats(x, m.weight, None, 1, k // 2, 1, groups, x2)
        y.backward(torch.ones_like(y))
    cases.append((name, run))


# convolutions: strip kernels (HBM-bound decoder / stem shapes), tile kernel 1x1 / 3x3 / grouped, transposed conv
conv_case("strip c32->32 @1024 (dec_l5)", 8, 32, 0, 1024, 1024, 32, 3)
conv_case("strip c64+64->64 @512 (dec_l4 conv1)", 8, 64, 64, 512, 512, 64, 3)
conv_case("strip c64->64 @512 (dec_l4 conv2)", 8, 64, 0, 512, 512, 64, 3)
conv_case("tile c128+256->128 @256 (dec_l3 conv1)", 8, 128, 256, 256, 256, 128, 3)
conv_case("tile c512+1024->512 @64 (dec_l1 conv1)", 8, 512, 1024, 64, 64, 512, 3)
conv_case("tile 1x1 c64->256 @256 (enc_l2 conv3)", 8, 64, 0, 256, 256, 256, 1)
conv_case("tile 1x1 c256->64 @256 (enc_l2 conv1)", 8, 256, 0, 256, 256, 64, 1)
conv_case("tile 1x1 c1024->256 @64 (enc_l4 conv1)", 8, 1024, 0, 64, 64, 256, 1)
conv_case("strip g2 c64->128 @256 (enc_l2 radix)", 8, 64, 0, 256, 256, 128, 3, 2)
conv_case("tile g2 c256->512 @64 (enc_l4 radix)", 8, 256, 0, 64, 64, 512, 3, 2)
xt = act(8, 64, 512, 512)
wt = nn.ConvTranspose2d(64, 32, 2, 2, bias=False).cuda()
wt.weight.data = wt.weight.data.contiguous(memory_format=CL)


def convt():
    y = ops.conv_transpose2x2(xt, wt.weight)
    y.backward(torch.ones_like(y))


cases.append(("convT c64->32 @512 (dec_l5 up)", convt))
xs = act(8, 3, 1024, 1024, grad=False)
ms = nn.Conv2d(3, 32, 3, 2, 1, bias=False).cuda()
ms.weight.data = ms.weight.data.contiguous(memory_format=CL)


def stem():
    y = ops.conv2d(xs, ms.weight, None, 2, 1, 1, 1)
    y.backward(torch.ones_like(y))


cases.append(("stem 3->32 s2 @1024", stem))
# BatchNorm streaming kernels (stats, train apply, bwd reduce, bwd apply) with and without residual
for name, c, h, res in (("bn lrelu c32 @1024", 32, 1024, False), ("bn relu+res c256 @256", 256, 256, True)):
    z, b = act(8, c, h, h), bn(c)
    r = act(8, c, h, h) if res else None

    def run(z=z, b=b, r=r, res=res):
        y = ops.batch_norm_act(z, b, ACT_RELU if res else ACT_LRELU, r)
        y.backward(torch.ones_like(y))
    cases.append((name, run))
# residual join whose gradient arrives in two parts (ops.fork): xv2_bn_bwd_reduce_du sums them while it streams
zf, bf_, rf = act(8, 256, 256, 256), bn(256), act(8, 256, 256, 256)
gf = [torch.ones(8, 256, 256, 256, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=CL) for _ in range(2)]


def bn_fork():
    ya, yb = ops.fork(ops.batch_norm_act(zf, bf_, ACT_RELU, rf))
    torch.autograd.backward([ya, yb], gf)
    ops.check_pending_addends()


cases.append(("bn relu+res c256 @256, two-part gradient (fork)", bn_fork))
# fused split attention (bn0 + relu folded in)
from xview2_b200.model.encoders import SplAtConv2d, _init_resnest
sp = SplAtConv2d(64, 1)
_init_resnest(sp)
sp = sp.cuda().train()
xsp = act(8, 64, 256, 256)


def splat():
    y = sp(xsp)
    y.backward(torch.ones_like(y))


cases.append(("SplAtConv2d c64 @256 (fused bn0+relu+attention)", splat))
# fused tail (last BN + LeakyReLU + head) and the loss
zt, bt = act(8, 32, 1024, 1024), bn(32)
hw_, hb_ = nn.Parameter(torch.randn(2, 32, 1, 1, device="cuda") * 0.1), nn.Parameter(torch.zeros(2, device="cuda"))
labels = torch.randint(0, 2, (8, 1024, 1024), device="cuda", dtype=torch.uint8)


def tail():
    logits = ops.bnact_head(ops.DeferredBNAct(zt, None, bt, ACT_LRELU), hw_, hb_)
    loss = ops.seg_loss(logits, labels, "focal+dice", False)
    loss.backward()


cases.append(("tail: bn+lrelu+head c32 @1024 + focal/dice loss", tail))
# pools
xp = act(8, 64, 512, 512)
cases.append(("maxpool 3x3 s2 c64 @512", lambda: ops.max_pool2d(xp, 3, 2, 1).backward(torch.ones(8, 64, 256, 256, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=CL))))
xa = act(8, 128, 256, 256)
cases.append(("avgpool 3x3 s2 c128 @256 (avd)", lambda: ops.avg_pool2d(xa, 3, 2, 1).backward(torch.ones(8, 128, 128, 128, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=CL))))
# loader / optimizer / post-process
tiles = torch.randint(0, 256, (8, 1024, 1024, 3), device="cuda", dtype=torch.uint8)
cases.append(("normalize_tiles 8 x 1024^2", lambda: ops.normalize_tiles(tiles)))
par = torch.zeros(8, 16)
par[:, 0:4] = torch.tensor([1024 / 1200, 1024 / 1200, 1200, 1200])
par[:, 10] = par[:, 12] = 1
par[:, 15] = 1
uni = torch.rand(8, 3)
cases.append(("augment_tiles 8 x 1024^2 -> 512^2 (zoom)", lambda: ops.augment_tiles(tiles, None, labels, par, uni)))
pflat = torch.randn(42_800_000, device="cuda")
st = [torch.zeros_like(pflat) for _ in range(3)]
cases.append(("adamw 42.8 M params", lambda: ops.adamw_step(pflat, st[0], st[1], st[2], 3e-4, 0.9, 0.999, 1e-8, 0.0, 1)))
post = (torch.rand(8, 1024, 1024, device="cuda") > 0.6).to(torch.uint8) * torch.randint(1, 5, (8, 1024, 1024), device="cuda", dtype=torch.uint8)
cases.append(("cc majority vote + dilate 8 x 1024^2", lambda: ops.dilate_square(ops.cc_majority_vote(post), 3)))
logits4 = torch.randn(8, 2, 1024, 1024, device="cuda").contiguous(memory_format=CL)
cnt = torch.zeros(3, dtype=torch.int64, device="cuda")
cases.append(("f1_update + argmax map 8 x 1024^2", lambda: ops.f1_update(logits4, labels, 2, cnt, torch.empty(8, 1024, 1024, dtype=torch.uint8, device="cuda"))))

if a.only:
    keys = [k for k in a.only.split(",") if k]
    cases = [(n, f) for n, f in cases if any(k in n for k in keys)]
if a.list:
    for n, _ in cases:
        print(n)
    sys.exit(0)
for _, fn in cases:  # warm-up: function attributes, allocator, packed weights
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for name, fn in cases:
    before = lib.launches()
    fn()
    torch.cuda.synchronize()
    print(f"{lib.launches() - before:3d} libxv2 launches  {name}", flush=True)
torch.cuda.profiler.stop()
