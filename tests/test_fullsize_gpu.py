"""GPU property tests at BASELINE.json's FULL sizes (8 x 1024^2 tiles), where the CPU oracle would take minutes: each
checks a size-independent invariant of the kernel instead of comparing with a reference.

  * translation equivariance of the strip convolution (bit-exact in the interior: same per-pixel arithmetic at another
    strip position / another CTA piece),
  * fused BN statistics == direct sums over the stored outputs,
  * weight-gradient linearity in dY and agreement of the two weight-gradient kernels on a shifted problem,
  * flip(flip(x)) == x, mean4 of four copies == the copy, F1 counter conservation, post-process idempotence,
  * loss additivity: the fused one-pass reduction over 8 tiles == the fp64 sum of per-tile reductions.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
CL = torch.channels_last


def _ops():
    from xview2_b200 import lib, ops
    lib.init(0)
    return ops


def _rand(shape, seed, dtype=torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, generator=g, device="cuda", dtype=torch.float32).to(dtype)


def test_strip_conv_translation_equivariance_fullsize():
    ops = _ops()
    n, c, h, w, k = 8, 32, 1024, 1024, 32
    x = _rand((n, c, h, w), 1).contiguous(memory_format=CL)
    wt = (_rand((k, c, 3, 3), 2, torch.float32) * (2.0 / (9 * c)) ** 0.5).contiguous(memory_format=CL)
    y, stats = ops.conv2d_stats(x, wt, None, 1, 1, 1, 1)
    assert stats is not None
    # shift the tile by (3 rows, 128 columns): interior outputs must be the SAME bits at the shifted position
    xs = torch.roll(x, shifts=(3, 128), dims=(2, 3)).contiguous(memory_format=CL)
    ys = ops.conv2d(xs, wt, None, 1, 1, 1, 1)
    a = y[:, :, 8:1000, 8:880]
    b = ys[:, :, 11:1003, 136:1008]
    assert torch.equal(a, b)
    # fused statistics == sums over the stored bf16 outputs
    o = y.double()
    ref = torch.cat((o.sum((0, 2, 3)), (o * o).sum((0, 2, 3))))
    assert float((stats - ref).abs().max() / ref.abs().max()) < 1e-5
    del o, ref
    # sanity against torch on one 64 x 64 window of one tile (fp32 conv of the same bf16 operands)
    import torch.nn.functional as F
    win = x[3:4, :, 500:566, 300:366].float()
    yr = F.conv2d(win, wt.to(torch.bfloat16).float())
    got = y[3:4, :, 501:565, 301:365].float()
    assert float((got - yr).abs().max() / yr.abs().max()) < 2e-2


def test_wgrad_linearity_and_shift_fullsize():
    ops = _ops()
    n, c, h, w, k = 8, 32, 1024, 1024, 32
    x = _rand((n, c, h, w), 3).contiguous(memory_format=CL)
    wt = torch.zeros(k, c, 3, 3, device="cuda").contiguous(memory_format=CL).requires_grad_(True)
    g1 = _rand((n, k, h, w), 4).contiguous(memory_format=CL)
    g1[:, :, :8] = 0   # no gradient near the top / bottom border: rolling by 5 rows then only permutes the pixel sum
    g1[:, :, -8:] = 0

    def wgrad(xx, gg):
        wt.grad = None
        ops.conv2d(xx, wt, None, 1, 1, 1, 1).backward(gg)
        return wt.grad.clone()

    d1 = wgrad(x, g1)
    d2 = wgrad(x, (g1.float() * 2).to(torch.bfloat16))  # exact doubling in bf16 -> exactly twice the gradient up to fp32 order
    assert float((d2 - 2 * d1).abs().max() / d1.abs().max()) < 1e-5
    # rolling x and dY together by whole rows permutes the pixel sum (other CTA pieces, other ring phases): same result
    d3 = wgrad(torch.roll(x, 5, 2).contiguous(memory_format=CL), torch.roll(g1, 5, 2).contiguous(memory_format=CL))
    assert float((d3 - d1).abs().max() / d1.abs().max()) < 1e-4


def test_flip_mean4_f1_postprocess_fullsize():
    ops = _ops()
    n, h, w = 8, 1024, 1024
    x = _rand((n, 32, h, w), 5).contiguous(memory_format=CL)
    assert torch.equal(ops.flip(ops.flip(x, [2, 3]), [2, 3]), x)
    assert torch.equal(ops.flip(ops.flip(x, [2]), [3]), ops.flip(x, [2, 3]))
    logits = _rand((n, 2, h, w), 6, torch.float32).contiguous(memory_format=CL)
    assert torch.equal(ops.mean4(logits, logits, logits, logits), logits)
    g = torch.Generator(device="cuda").manual_seed(7)
    labels = torch.randint(0, 2, (n, h, w), generator=g, device="cuda", dtype=torch.uint8)
    counters = torch.zeros(3, dtype=torch.int64, device="cuda")
    pred = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    ops.f1_update(logits, labels, 2, counters, pred)
    tp, fp, fn = (int(v) for v in counters)
    assert tp + fn == int((labels == 1).sum()) and tp + fp == int((pred == 1).sum())
    assert torch.equal(pred, logits.argmax(1).to(torch.uint8))  # bit-exact argmax label map (lowest index wins ties)
    dmg = _rand((n, 4, h, w), 8, torch.float32).contiguous(memory_format=CL)
    pre_map, post_map = ops.post_process(logits, dmg)
    assert int(post_map.max()) <= 4 and bool(((post_map > 0) == (pre_map > 0)).all())
    loc = torch.sigmoid(logits[:, 1])
    keep = (loc > 0.3) | ((loc > 0.1) & ((dmg.argmax(1) + 1) > 1))
    band = ((loc - 0.3).abs() < 1e-6) | ((loc - 0.1).abs() < 1e-6)  # threshold ties may differ by one ulp of the sigmoid
    assert bool(((pre_map > 0) == keep)[~band].all())


def test_loss_additivity_fullsize():
    ops = _ops()
    from xview2_b200 import lib
    n, h, w = 8, 1024, 1024
    logits = _rand((n, 2, h, w), 9, torch.float32).contiguous(memory_format=CL)
    g = torch.Generator(device="cuda").manual_seed(10)
    labels = torch.randint(0, 2, (n, h, w), generator=g, device="cuda", dtype=torch.uint8)
    whole = torch.zeros(9, dtype=torch.float64, device="cuda")
    lib.call("xv2_loss_partials", lib.ptr(logits), lib.ptr(labels), n, h, w, 2, 1, 0, lib.ptr(whole))
    parts = torch.zeros(9, dtype=torch.float64, device="cuda")
    for i in range(n):
        li, yi = logits[i:i + 1].contiguous(memory_format=CL), labels[i:i + 1].contiguous()
        lib.call("xv2_loss_partials", lib.ptr(li), lib.ptr(yi), 1, h, w, 2, 1, 0, lib.ptr(parts))
    assert float((whole - parts).abs().max() / whole.abs().max()) < 1e-9
    assert int(whole[-1]) == n * h * w
    # and the fused value equals the plain-torch focal + dice on the same logits (MONAI 0.4 definitions, loss.py:11-13)
    loss = ops.seg_loss(logits, labels, "focal+dice", False)
    p = torch.softmax(logits.double(), 1)
    t1 = labels.double()
    dice = 1 - (2 * (p[:, 1] * t1).sum() + 1e-5) / (t1.sum() + p[:, 1].sum() + 1e-5)
    logpt = torch.log_softmax(logits.double(), 1).gather(1, labels.long().unsqueeze(1)).squeeze(1)
    focal = (-(1 - logpt.exp()) ** 2 * logpt).mean()
    assert abs(float(loss) - float(dice + focal)) < 1e-5
