"""GPU block tests (-m gpu): forward AND backward of the building blocks of the hot path through libxv2 against the CPU
oracle on identical, well-conditioned inputs (random activations, random upstream gradients -- no 100-layer error
amplification), fp32 path.  This is where gradient parity is held tight: 1e-4 forward, 1e-3 gradients (max-norm
relative), train-mode BN semantics included.  Whole-network parity lives in tests/test_model_gpu.py.
"""
import argparse

import pytest
import torch

from oracle import functional as OF
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def _state(module, seed):
    shapes = {k: (tuple(v.shape), v.dtype) for k, v in module.state_dict().items()}
    st = OF.deterministic_state(shapes, seed)
    module.load_state_dict(st, strict=True)
    return st


def _leaves(st, prefix=""):
    return {prefix + k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v.clone())
            for k, v in st.items()}


def _rand(shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def _compare_grads(module, P, prefix, tol=1e-3, skip=()):
    for k, p in module.named_parameters():
        if any(k.endswith(s) for s in skip):
            continue
        ref = P[prefix + k].grad
        assert ref is not None and p.grad is not None, k
        assert rel_err(p.grad, ref) < tol, (k, rel_err(p.grad, ref))


@pytest.mark.parametrize("stride,first,down", [(1, False, False), (2, True, True), (1, False, True)])
def test_resnest_bottleneck(stride, first, down):
    from xview2_b200 import ops
    from xview2_b200.model.encoders import SplAtBottleneck, _init_resnest
    inpl, planes = (256 if not down else 128), 64
    blk = SplAtBottleneck(inpl, planes, stride, 1, is_first=first, with_down=down, down_pool=stride)
    st = _state(blk, 3)
    blk = blk.cuda().train()
    x = _rand((4, inpl, 16, 16), 1)
    gy = _rand((4, planes * 4, 16 // stride, 16 // stride), 2)
    xi = ops.nhwc(x.cuda()).requires_grad_(True)
    out = blk(xi)
    out.backward(ops.nhwc(gy.cuda()))
    P = _leaves(st, "b.")
    xr = x.clone().requires_grad_(True)
    ref = OF._resnest_block(P, "b", xr, True, stride, 1, first, down)
    ref.backward(gy)
    assert rel_err(out, ref) < 1e-4
    assert rel_err(xi.grad, xr.grad) < 1e-3
    _compare_grads(blk, P, "b.", skip=("conv2.fc1.bias",))
    for k in ("bn1.running_mean", "bn3.running_var", "conv2.bn1.running_var"):
        assert rel_err(blk.state_dict()[k], P["b." + k]) < 1e-4


def test_resnet_bottleneck():
    from xview2_b200 import ops
    from xview2_b200.model.encoders import Bottleneck
    blk = Bottleneck(64, 32, 2, 1, True)
    st = _state(blk, 4)
    blk = blk.cuda().train()
    x, gy = _rand((4, 64, 16, 16), 1), _rand((4, 128, 8, 8), 2)
    xi = ops.nhwc(x.cuda()).requires_grad_(True)
    out = blk(xi)
    out.backward(ops.nhwc(gy.cuda()))
    P = _leaves(st, "b.")
    xr = x.clone().requires_grad_(True)
    ref = OF._resnet_block(P, "b", xr, True, 2, 1, True)
    ref.backward(gy)
    assert rel_err(out, ref) < 1e-4 and rel_err(xi.grad, xr.grad) < 1e-3
    _compare_grads(blk, P, "b.")


@pytest.mark.parametrize("attention,skip_c", [(False, 64), (True, 64), (False, 0)])
def test_upsample_block(attention, skip_c):
    from xview2_b200 import ops
    from xview2_b200.model.layers import UpsampleBlock
    blk = UpsampleBlock(128, 32, skip_c, attention, False)
    st = _state(blk, 5)
    blk = blk.cuda().train()
    x, gy = _rand((4, 128, 8, 8), 1), _rand((4, 32, 16, 16), 2)
    sk = _rand((4, skip_c, 16, 16), 3) if skip_c else None
    xi = ops.nhwc(x.cuda()).requires_grad_(True)
    si = ops.nhwc(sk.cuda()).requires_grad_(True) if skip_c else None
    out = blk(xi, si)
    out.backward(ops.nhwc(gy.cuda()))
    P = _leaves(st, "u.")
    xr = x.clone().requires_grad_(True)
    sr = sk.clone().requires_grad_(True) if skip_c else None
    ref = OF.upsample_block(P, "u", xr, sr, True, attention)
    ref.backward(gy)
    assert rel_err(out, ref) < 1e-4 and rel_err(xi.grad, xr.grad) < 1e-3
    if skip_c:
        assert rel_err(si.grad, sr.grad) < 1e-3
    _compare_grads(blk, P, "u.")


def test_fusion_block_and_loss_chain():
    """FusionBlock (layers.py:103-116) over two ConvLayer stages + 4-class head + masked focal+dice loss."""
    from xview2_b200 import ops
    from xview2_b200.model.layers import ConvLayer, FusionBlock, OutputBlock
    from xview2_b200.model.loss import Loss
    pre_c, post_c = ConvLayer(16, 32), ConvLayer(16, 32)
    fb = FusionBlock(pre_c, post_c, 32)
    head = OutputBlock(64, 4, False)
    st, sh = _state(fb, 6), _state(head, 7)
    fb, head = fb.cuda().train(), head.cuda().train()
    a, b = _rand((4, 16, 16, 16), 1), _rand((4, 16, 16, 16), 2)
    y = torch.randint(0, 5, (4, 16, 16), generator=torch.Generator().manual_seed(3), dtype=torch.uint8)
    ai, bi = ops.nhwc(a.cuda()).requires_grad_(True), ops.nhwc(b.cuda()).requires_grad_(True)
    pre, post = fb(ai, bi)
    logits = head(pre, post)
    ns = argparse.Namespace(loss_str="focal+dice", type="post")
    loss = Loss(ns)(logits, y.cuda())
    loss.backward()
    P = {**_leaves(st, "f."), **_leaves(sh, "h.")}
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    rpre, rpost = OF._fusion(P, "f", OF.conv_layer(P, "f.pre_conv", ar, True), OF.conv_layer(P, "f.post_conv", br, True), True)
    rl = torch.nn.functional.conv2d(torch.cat((rpre, rpost), 1), P["h.conv.weight"], P["h.conv.bias"])
    rloss = OF.loss_forward(rl, y, "focal+dice", True)
    rloss.backward()
    assert rel_err(logits, rl) < 1e-4
    assert abs(float(loss.detach()) - float(rloss.detach())) < 1e-5
    assert rel_err(ai.grad, ar.grad) < 1e-3 and rel_err(bi.grad, br.grad) < 1e-3
    _compare_grads(fb, P, "f.")
    _compare_grads(head, P, "h.")


@pytest.mark.parametrize("c,ncls,hw", [(32, 2, (24, 40)), (64, 4, (16, 24)), (32, 1, (9, 16)), (64, 3, (8, 8))])
def test_fused_bnact_head_tail(c, ncls, hw):
    """Last ConvLayer's train-mode BatchNorm + LeakyReLU fused with the 1x1 head (csrc/fused_tail.cu) against plain PyTorch fp32
    on the same bf16 conv output: logits, dz, BatchNorm and head parameter gradients, running statistics."""
    import torch.nn.functional as F

    from xview2_b200 import ops
    from xview2_b200.lib import ACT_LRELU
    g = torch.Generator().manual_seed(31)
    z = (torch.randn(3, c, *hw, generator=g) * 1.5 + 0.3).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    z.requires_grad_(True)
    bn = torch.nn.BatchNorm2d(c).cuda().train()
    with torch.no_grad():
        bn.weight.copy_(0.5 + torch.rand(c, generator=g))
        bn.bias.copy_(0.2 * torch.randn(c, generator=g))
    hw_ = torch.nn.Parameter((torch.randn(ncls, c, 1, 1, generator=g) * 0.3).cuda())
    hb_ = torch.nn.Parameter((torch.randn(ncls, generator=g) * 0.1).cuda())
    logits = ops.bnact_head(ops.DeferredBNAct(z, None, bn, ACT_LRELU), hw_, hb_)
    gl = torch.randn(*logits.shape, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    logits.backward(gl)

    zr = z.detach().float().requires_grad_(True)
    bnr = torch.nn.BatchNorm2d(c).cuda().train()
    with torch.no_grad():
        bnr.weight.copy_(bn.weight)
        bnr.bias.copy_(bn.bias)
    wr, br = hw_.detach().clone().requires_grad_(True), hb_.detach().clone().requires_grad_(True)
    y = F.leaky_relu(bnr(zr), 0.01)
    ref = F.conv2d(y, wr, br)
    ref.backward(gl)

    def rel(a, b):
        a, b = a.detach().double().cpu(), b.detach().double().cpu()
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
    assert logits.dtype == torch.float32 and rel(logits, ref) < 6e-3       # y is rounded to bf16 like the unfused path stores it
    assert rel(z.grad, zr.grad) < 1e-2                                      # dz is stored in bf16
    assert rel(bn.weight.grad, bnr.weight.grad) < 5e-3 and rel(bn.bias.grad, bnr.bias.grad) < 5e-3
    assert rel(hw_.grad, wr.grad) < 5e-3 and rel(hb_.grad, br.grad) < 1e-4
    assert rel(bn.running_mean, bnr.running_mean) < 1e-4 and rel(bn.running_var, bnr.running_var) < 1e-4
    assert int(bn.num_batches_tracked) == 1


@pytest.mark.parametrize("channels,hw,n", [(64, (16, 16), 4), (128, (8, 16), 3), (512, (8, 8), 2)])
def test_fused_bn_split_attention(channels, hw, n):
    """SplAtConv2d with bn0 + ReLU folded into the split-attention kernels (csrc/splat_fused.cu: the 2C-channel activation is
    never written, the BatchNorm backward reductions come from per-image partial sums) against the oracle's SplAt in fp32."""
    from xview2_b200 import ops
    from xview2_b200.model.encoders import SplAtConv2d, _init_resnest
    torch.manual_seed(3)
    mod = SplAtConv2d(channels, 1)
    _init_resnest(mod)
    g = torch.Generator().manual_seed(17)
    with torch.no_grad():
        mod.conv.weight.copy_(mod.conv.weight.to(torch.bfloat16).float())  # both sides see the same (bf16-exact) conv weights
        for bn in (mod.bn0, mod.bn1):
            bn.weight.copy_(0.5 + torch.rand(bn.num_features, generator=g))
            bn.bias.copy_(0.2 * torch.randn(bn.num_features, generator=g))
        mod.fc2.weight.mul_(0.25)
    mod = mod.cuda().train()
    x = torch.randn(n, channels, *hw, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    calls = []
    orig = ops.call

    def spy(fn, *a, **kw):
        calls.append(fn)
        return orig(fn, *a, **kw)

    def run(fused):
        import copy
        m2 = copy.deepcopy(mod)
        xx = x.detach().clone().requires_grad_(True)
        ops.FUSE_SPLAT_BN = fused
        try:
            o = m2(xx)
            o.backward(gy)
        finally:
            ops.FUSE_SPLAT_BN = True
        return m2, xx, o

    gy = torch.randn(n, channels, *hw, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ops.call = spy
    try:
        mod_f, x_f, out = run(True)
    finally:
        ops.call = orig
    assert ("xv2_splat_bn_gap_fin" in calls and "xv2_splat_fc_bwd_fused" in calls and "xv2_splat_bn_bwd_apply" in calls and
            "xv2_bn_train_apply" not in calls and "xv2_bn_finalize" not in calls), calls
    mod_u, x_u, out_u = run(False)  # the unfused chain on the same inputs: the yard-stick for what bf16 storage costs

    P = {"m." + k: v.detach().float().cpu().clone() for k, v in mod.state_dict().items()}
    for k in list(P):
        if "running_mean" in k:
            P[k].zero_()
        elif "running_var" in k:
            P[k].fill_(1.0)
        elif "num_batches" in k:
            P[k] = torch.zeros((), dtype=torch.int64)
    leaves = {k: v.requires_grad_(v.is_floating_point() and "running_" not in k) for k, v in P.items()}
    xr = x.detach().float().cpu().requires_grad_(True)
    ref = OF._splat(leaves, "m", xr, True, 1)
    ref.backward(gy.float().cpu())
    assert rel_err(out, ref) < max(2e-2, 1.5 * rel_err(out_u, ref))
    assert rel_err(x_f.grad, xr.grad) < max(3e-2, 1.5 * rel_err(x_u.grad, xr.grad)), (rel_err(x_f.grad, xr.grad), rel_err(x_u.grad, xr.grad))
    named, named_u = dict(mod_f.named_parameters()), dict(mod_u.named_parameters())
    for k in ("bn0.weight", "bn0.bias", "fc1.weight", "fc2.weight", "fc2.bias", "bn1.weight", "bn1.bias", "conv.weight"):
        e_f, e_u = rel_err(named[k].grad, leaves["m." + k].grad), rel_err(named_u[k].grad, leaves["m." + k].grad)
        assert e_f < max(3e-2, 1.5 * e_u), (k, e_f, e_u)
    sd = mod_f.state_dict()
    assert rel_err(sd["bn0.running_mean"], leaves["m.bn0.running_mean"]) < 1e-2
    assert rel_err(sd["bn0.running_var"], leaves["m.bn0.running_var"]) < 1e-2
    assert int(sd["bn0.num_batches_tracked"]) == 1 and int(sd["bn1.num_batches_tracked"]) == 1


@pytest.mark.parametrize("kind", ["resnest", "resnet"])
def test_fork_gradients_match_autograd_accumulation(kind, monkeypatch):
    """bf16 identity-shortcut blocks: ops.fork hands the shortcut's gradient part to the producer's BatchNorm backward
    (xv2_bn_bwd_reduce_du sums dy + dy2 while it streams) -- same gradients as autograd's own accumulation pass."""
    from xview2_b200 import ops
    from xview2_b200.model.encoders import Bottleneck, SplAtBottleneck, _BlockList
    if kind == "resnest":
        blocks = [SplAtBottleneck(128, 64, 1, 1, False, True, 1)] + [SplAtBottleneck(256, 64, 1, 1, False, False, 1) for _ in range(2)]
    else:
        blocks = [Bottleneck(128, 64, 1, 1, True)] + [Bottleneck(256, 64, 1, 1, False) for _ in range(2)]
    stage = _BlockList(blocks)
    _state(stage, 11)
    stage = stage.cuda().train()
    # samples of clearly different scale: keeps the split attention's BatchNorm over the n = 4 pooled vectors well-conditioned
    x = ops.nhwc((_rand((4, 128, 32, 32), 1) * torch.tensor([0.5, 1.0, 2.0, 4.0]).view(4, 1, 1, 1)).cuda().to(torch.bfloat16))
    gy = ops.nhwc(_rand((4, 256, 32, 32), 2).cuda().to(torch.bfloat16))
    seen = []
    real_call = ops.call

    def spy(name, *args, **kw):
        if name == "xv2_bn_bwd_reduce_du":
            seen.append(args[1] is not None)  # dy2 present?
        return real_call(name, *args, **kw)

    monkeypatch.setattr(ops, "call", spy)
    results = []
    for fork in (True, False, False):
        monkeypatch.setattr(ops, "FORK_GRADS", fork)
        seen.clear()
        stage.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_(True)
        stage(xi).backward(gy)
        ops.check_pending_addends()
        results.append((xi.grad.float().clone(), {k: p.grad.float().clone() for k, p in stage.named_parameters()}, list(seen)))
    (gx_f, gp_f, seen_f), (gx_a, gp_a, seen_a), (gx_b, gp_b, _) = results
    assert sum(seen_f) == 2 and sum(seen_a) == 0, (seen_f, seen_a)  # two identity blocks -> two parked parts, all consumed
    assert len(seen_f) == 3 and len(seen_a) == 3                    # every bn3 (residual join) takes the du path

    def l2(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

    # Both stages are run-to-run reproducible in forward and data gradient (BatchNorm statistics, split-attention pooling and
    # partial sums are accumulated in fp64; the fp32 red.global adds of the split-K weight-gradient kernels only touch parameter
    # gradients), so the fork route must be BIT-identical to autograd's accumulation for the data gradient ...
    assert torch.equal(gx_a, gx_b), "two identical runs differ"
    assert torch.equal(gx_f, gx_a), l2(gx_f, gx_a)
    # ... and within the summation-order noise of the weight-gradient atomics for the parameters
    for k in gp_f:
        if k.endswith("conv2.fc1.bias"):  # analytically zero (bias in front of a BatchNorm): only noise to compare
            continue
        assert rel_err(gp_f[k], gp_a[k]) < 1e-4, (k, rel_err(gp_f[k], gp_a[k]), rel_err(gp_b[k], gp_a[k]))


def test_flat_gradient_slots_receive_fc_and_head_gradients():
    """With optim.FlatParams the split-attention FC chain and the fused BN + head kernels ADD their parameter gradients into
    the parameters' own slots of the flat gradient buffer (no autograd accumulation launches): same values as the gradients
    autograd returns without the flat buffer, and a second backward without zero_grad doubles them."""
    import copy

    from torch import nn

    from xview2_b200 import ops
    from xview2_b200.lib import ACT_LRELU
    from xview2_b200.model.encoders import SplAtConv2d, _init_resnest
    from xview2_b200.optim import FlatParams

    class Tiny(nn.Module):
        def __init__(self):
            super().__init__()
            self.sp = SplAtConv2d(64, 1)
            self.bn = nn.BatchNorm2d(64)
            self.head = nn.Conv2d(64, 2, 1)

        def forward(self, x):
            y = self.sp(x)
            return ops.bnact_head(ops.DeferredBNAct(y, None, self.bn, ACT_LRELU), self.head.weight, self.head.bias)

    torch.manual_seed(9)
    net = Tiny()
    _init_resnest(net.sp)
    net = net.cuda().train()
    g = torch.Generator().manual_seed(4)
    x = ops.nhwc((torch.randn(4, 64, 32, 32, generator=g) * torch.tensor([0.5, 1.0, 2.0, 4.0]).view(4, 1, 1, 1)).cuda().to(torch.bfloat16))
    gl = torch.randn(4, 2, 32, 32, generator=g).cuda().contiguous(memory_format=torch.channels_last)

    def grads(flat, passes=1):
        m = copy.deepcopy(net)
        fp = FlatParams(m) if flat else None
        if fp is not None:
            fp.zero_grad()
        for _ in range(passes):
            m(x).backward(gl)
        ops.check_pending_addends()
        torch.cuda.synchronize()
        return {k: p.grad.detach().float().reshape(-1).clone() for k, p in m.named_parameters()}

    def l2(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

    ga, gb = grads(False), grads(False)  # run-to-run noise of the plain path (fp32 atomics of the weight-gradient kernels)
    gf, gf2 = grads(True), grads(True, passes=2)
    for k in ga:
        if k.endswith("fc1.bias"):  # analytically zero (bias in front of a BatchNorm)
            continue
        tol = max(4 * l2(gb[k], ga[k]), 1e-4)  # the forward is reproducible; only the weight-gradient atomics add order noise
        assert l2(gf[k], ga[k]) <= tol, (k, l2(gf[k], ga[k]), tol)
        assert l2(gf2[k], 2 * ga[k]) <= 2 * tol, (k, l2(gf2[k], 2 * ga[k]), tol)
