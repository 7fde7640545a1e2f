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
