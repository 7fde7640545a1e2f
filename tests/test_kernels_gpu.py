"""GPU unit tests (-m gpu): every libxv2 operator against a plain PyTorch fp32 reference of the same op.

Tolerances: fp32 path 1e-4 relative (accumulation order only); bf16 path 2e-2 relative to the tensor's max
(bf16 has 8 mantissa bits; inputs are rounded to bf16 first so only accumulation/rounding of outputs differ).
Integer / label outputs (argmax maps, F1 counters, post-process) are compared bit-exactly.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CL = torch.channels_last


def _ops():
    from xview2_b200 import ops
    return ops


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def rnd(*shape, dtype=torch.float32, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(*shape, generator=g) * scale
    return t.to(dtype).cuda()


def tol(dtype):
    return 1e-4 if dtype == torch.float32 else 2e-2


CONV_CASES = [
    # n, c, h, w, k, r, stride, pad, dil, groups
    (2, 3, 32, 32, 32, 3, 2, 1, 1, 1),      # ResNeSt stem conv1
    (2, 3, 32, 32, 64, 7, 2, 3, 1, 1),      # ResNet stem
    (2, 32, 16, 16, 64, 3, 1, 1, 1, 1),
    (2, 64, 16, 16, 128, 3, 1, 1, 1, 2),    # SplAt grouped conv
    (2, 64, 16, 16, 64, 3, 2, 1, 1, 1),     # ResNet strided 3x3
    (2, 128, 8, 8, 256, 1, 1, 0, 1, 1),
    (2, 128, 16, 16, 256, 1, 2, 0, 1, 1),   # ResNet strided 1x1 downsample
    (1, 64, 16, 16, 64, 3, 1, 2, 2, 1),     # dilated
    (3, 40, 9, 11, 24, 3, 1, 1, 1, 1),      # ragged sizes
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_simt_fwd_bwd(case, dtype):
    ops = _ops()
    n, c, h, w, k, r, stride, pad, dil, groups = case
    x = rnd(n, c, h, w, dtype=dtype, seed=1).contiguous(memory_format=CL).requires_grad_(True)
    wt = rnd(k, c // groups, r, r, seed=2, scale=(2.0 / (c // groups * r * r)) ** 0.5).contiguous(memory_format=CL).requires_grad_(True)
    b = rnd(k, seed=3).requires_grad_(True)
    old = ops.USE_TENSOR_CORES
    ops.USE_TENSOR_CORES = False
    try:
        y = ops.conv2d(x, wt, b, stride, pad, dil, groups)
        gy = rnd(*y.shape, dtype=dtype, seed=4)
        y.backward(gy)
    finally:
        ops.USE_TENSOR_CORES = old
    xr = x.detach().float().requires_grad_(True)
    wr = (wt.detach().to(dtype).float() if dtype != torch.float32 else wt.detach().clone()).requires_grad_(True)
    br = b.detach().clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, br, stride, pad, dil, groups)
    yr.backward(gy.float())
    t = tol(dtype)
    assert rel(y, yr) < t
    assert rel(x.grad, xr.grad) < t
    assert rel(wt.grad, wr.grad) < t
    assert rel(b.grad, br.grad) < t


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_transpose_simt(dtype):
    ops = _ops()
    x = rnd(2, 32, 8, 8, dtype=dtype, seed=1).contiguous(memory_format=CL).requires_grad_(True)
    wt = rnd(32, 16, 2, 2, seed=2, scale=0.2).contiguous(memory_format=CL).requires_grad_(True)
    old = ops.USE_TENSOR_CORES
    ops.USE_TENSOR_CORES = False
    try:
        y = ops.conv_transpose2x2(x, wt)
        gy = rnd(*y.shape, dtype=dtype, seed=4)
        y.backward(gy)
    finally:
        ops.USE_TENSOR_CORES = old
    xr = x.detach().float().requires_grad_(True)
    wr = (wt.detach().to(dtype).float()).requires_grad_(True)
    yr = F.conv_transpose2d(xr, wr, None, 2)
    yr.backward(gy.float())
    t = tol(dtype)
    assert rel(y, yr) < t and rel(x.grad, xr.grad) < t and rel(wt.grad, wr.grad) < t


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("act,with_res", [(0, False), (1, False), (2, False), (1, True)])
@pytest.mark.parametrize("c", [32, 20])
def test_batch_norm_act(dtype, training, act, with_res, c):
    ops = _ops()
    bn = torch.nn.BatchNorm2d(c).cuda()
    bn.weight.data = rnd(c, seed=5) * 0.2 + 1
    bn.bias.data = rnd(c, seed=6) * 0.2
    bn.running_mean.data = rnd(c, seed=7) * 0.1
    bn.running_var.data = rnd(c, seed=8).abs() + 0.5
    bn.train(training)
    ref = torch.nn.BatchNorm2d(c).cuda()
    ref.load_state_dict(bn.state_dict())
    ref.train(training)
    x = rnd(4, c, 12, 10, dtype=dtype, seed=1).contiguous(memory_format=CL).requires_grad_(True)
    res = rnd(4, c, 12, 10, dtype=dtype, seed=2).contiguous(memory_format=CL).requires_grad_(True) if with_res else None
    y = ops.batch_norm_act(x, bn, act, res)
    gy = rnd(*y.shape, dtype=dtype, seed=3)
    y.backward(gy)
    xr = x.detach().float().requires_grad_(True)
    rr = res.detach().float().requires_grad_(True) if with_res else None
    u = ref(xr)
    if with_res:
        u = u + rr
    yr = u if act == 0 else (F.relu(u) if act == 1 else F.leaky_relu(u, 0.01))
    yr.backward(gy.float())
    t = tol(dtype)
    assert rel(y, yr) < t
    assert rel(x.grad, xr.grad) < t * 2
    assert rel(bn.weight.grad, ref.weight.grad) < t * 2 and rel(bn.bias.grad, ref.bias.grad) < t * 2
    if with_res:
        assert rel(res.grad, rr.grad) < t
    if training:
        assert rel(bn.running_mean, ref.running_mean) < 1e-3 and rel(bn.running_var, ref.running_var) < 1e-3
        assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked) == 1


@pytest.mark.parametrize("k,hw", [(32, (64, 96)), (64, (33, 47))])
def test_stem_conv_direct(k, hw):
    """Cin = 3 stride-2 stem conv (ResNeSt conv1[0]) runs the direct kernel: forward + weight gradient vs torch fp32."""
    ops = _ops()
    h, w = hw
    x = rnd(2, 3, h, w, dtype=torch.bfloat16, seed=1).contiguous(memory_format=CL)
    wt = (rnd(k, 3, 3, 3, seed=2) * 0.3).contiguous(memory_format=CL).requires_grad_(True)
    y = ops.conv2d(x, wt, None, 2, 1, 1, 1)
    gy = rnd(*y.shape, dtype=torch.bfloat16, seed=3).contiguous(memory_format=CL)
    y.backward(gy)
    wr = wt.detach().clone().requires_grad_(True)
    yr = F.conv2d(x.float(), wr, None, 2, 1)
    yr.backward(gy.float())
    assert y.shape == yr.shape
    assert rel(y, yr) < 1e-2 and rel(wt.grad, wr.grad) < 1e-3


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("act,with_res,c,hw", [(2, False, 32, (100, 83)), (1, True, 256, (50, 41)), (0, False, 2048, (17, 16))])
def test_batch_norm_stream(training, act, with_res, c, hw):
    """Large bf16 tensors take the shared-memory-staged streaming kernels (bn_stream.cu); ragged last chunk included."""
    ops = _ops()
    dtype = torch.bfloat16
    bn = torch.nn.BatchNorm2d(c).cuda()
    bn.weight.data = rnd(c, seed=5) * 0.2 + 1
    bn.bias.data = rnd(c, seed=6) * 0.2
    bn.running_mean.data = rnd(c, seed=7) * 0.1
    bn.running_var.data = rnd(c, seed=8).abs() + 0.5
    bn.train(training)
    ref = torch.nn.BatchNorm2d(c).cuda()
    ref.load_state_dict(bn.state_dict())
    ref.train(training)
    h, w = hw
    x = rnd(4, c, h, w, dtype=dtype, seed=1).contiguous(memory_format=CL).requires_grad_(True)
    assert x.numel() >= 1 << 20
    res = rnd(4, c, h, w, dtype=dtype, seed=2).contiguous(memory_format=CL).requires_grad_(True) if with_res else None
    y = ops.batch_norm_act(x, bn, act, res)
    gy = rnd(*y.shape, dtype=dtype, seed=3).contiguous(memory_format=CL)
    y.backward(gy)
    xr = x.detach().float().requires_grad_(True)
    rr = res.detach().float().requires_grad_(True) if with_res else None
    u = ref(xr)
    if with_res:
        u = u + rr
    yr = u if act == 0 else (F.relu(u) if act == 1 else F.leaky_relu(u, 0.01))
    yr.backward(gy.float())
    t = tol(dtype)
    assert rel(y, yr) < t
    assert rel(x.grad, xr.grad) < t * 2
    assert rel(bn.weight.grad, ref.weight.grad) < t * 2 and rel(bn.bias.grad, ref.bias.grad) < t * 2
    if with_res:
        assert rel(res.grad, rr.grad) < t
    if training:
        assert rel(bn.running_mean, ref.running_mean) < 1e-3 and rel(bn.running_var, ref.running_var) < 1e-3


@pytest.mark.parametrize("hw", [(15, 18), (16, 20)])  # even sizes take the 2x2-quad backward kernels (3x3 / 2 / 1 pools)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pools(dtype, hw):
    ops = _ops()
    t = tol(dtype)
    x = rnd(2, 16, hw[0], hw[1], dtype=dtype, seed=1).contiguous(memory_format=CL).requires_grad_(True)
    # reference in plain NCHW: torch 2.11's CUDA channels-last avg_pool2d BACKWARD is wrong for padded non-square
    # inputs (it disagrees with torch's own CPU and NCHW-CUDA results; tools/dbg_pool.py shows it)
    xr = x.detach().float().contiguous().requires_grad_(True)
    for fn, rf in [
        (lambda a: ops.max_pool2d(a, 3, 2, 1), lambda a: F.max_pool2d(a, 3, 2, 1)),
        (lambda a: ops.avg_pool2d(a, 3, 2, 1), lambda a: F.avg_pool2d(a, 3, 2, 1)),
        (lambda a: ops.avg_pool2d(a, 3, 1, 1), lambda a: F.avg_pool2d(a, 3, 1, 1)),
        (lambda a: ops.avg_pool2d(a, 2, 2, 0, True, False), lambda a: F.avg_pool2d(a, 2, 2, 0, True, False)),
    ]:
        x.grad = None
        xr.grad = None
        y, yr = fn(x), rf(xr)
        assert y.shape == yr.shape
        gy = rnd(*y.shape, dtype=dtype, seed=9)
        y.backward(gy)
        yr.backward(gy.float())
        assert rel(y, yr) < t and rel(x.grad, xr.grad) < t


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("training", [True, False])
def test_split_attention(dtype, training):
    ops = _ops()
    c, n = 32, 3
    fc1 = torch.nn.Conv2d(c, 32, 1).cuda()
    bn1 = torch.nn.BatchNorm2d(32).cuda()
    fc2 = torch.nn.Conv2d(32, 2 * c, 1).cuda()
    bn1.running_var.data.fill_(0.7)
    bn1.train(training)
    x = rnd(n, 2 * c, 6, 5, dtype=dtype, seed=1).contiguous(memory_format=CL).requires_grad_(True)
    out = ops.split_attention(x, fc1, bn1, fc2)
    gy = rnd(*out.shape, dtype=dtype, seed=2)
    out.backward(gy)
    got = {k: p.grad.clone() for k, p in [("w1", fc1.weight), ("b1", fc1.bias), ("g", bn1.weight), ("b", bn1.bias),
                                          ("w2", fc2.weight), ("b2", fc2.bias)]}
    for p in (fc1.weight, fc1.bias, bn1.weight, bn1.bias, fc2.weight, fc2.bias):
        p.grad = None
    bn1.running_mean.data.zero_()
    bn1.running_var.data.fill_(0.7)
    xr = x.detach().float().requires_grad_(True)
    x0, x1 = xr[:, :c], xr[:, c:]
    gap = F.adaptive_avg_pool2d(x0 + x1, 1)
    a = fc2(F.relu(bn1(fc1(gap))))
    a = torch.softmax(a.view(n, 1, 2, c).transpose(1, 2), 1).reshape(n, 2 * c, 1, 1)
    yr = a[:, :c] * x0 + a[:, c:] * x1
    yr.backward(gy.float())
    t = tol(dtype)
    assert rel(out, yr) < t
    assert rel(x.grad, xr.grad) < t * 2
    assert rel(got["w2"], fc2.weight.grad) < t * 3 and rel(got["b2"], fc2.bias.grad) < t * 3
    assert rel(got["w1"], fc1.weight.grad) < t * 3
    assert rel(got["g"], bn1.weight.grad) < t * 3 and rel(got["b"], bn1.bias.grad) < t * 3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_elementwise(dtype):
    ops = _ops()
    t = tol(dtype)
    a = rnd(2, 16, 7, 9, dtype=dtype, seed=1).contiguous(memory_format=CL).requires_grad_(True)
    b = rnd(2, 16, 7, 9, dtype=dtype, seed=2).contiguous(memory_format=CL).requires_grad_(True)
    y = ops.add_act(a, b, 1)
    gy = rnd(*y.shape, dtype=dtype, seed=3)
    y.backward(gy)
    ar, br = a.detach().float().requires_grad_(True), b.detach().float().requires_grad_(True)
    yr = F.relu(ar + br)
    yr.backward(gy.float())
    assert rel(y, yr) < t and rel(a.grad, ar.grad) < t and rel(b.grad, br.grad) < t
    # attention gate
    skip = rnd(2, 16, 7, 9, dtype=dtype, seed=4).contiguous(memory_format=CL).requires_grad_(True)
    psi = rnd(2, 1, 7, 9, dtype=dtype, seed=5).contiguous(memory_format=CL).requires_grad_(True)
    o = ops.gate(skip, psi)
    o.backward(gy)
    sr, pr = skip.detach().float().requires_grad_(True), psi.detach().float().requires_grad_(True)
    orf = sr * torch.sigmoid(pr)
    orf.backward(gy.float())
    assert rel(o, orf) < t and rel(skip.grad, sr.grad) < t and rel(psi.grad, pr.grad) < t * 2
    # flips are exact
    for dims in ([2], [3], [2, 3]):
        assert torch.equal(ops.flip(a.detach(), dims), torch.flip(a.detach(), dims))
    # cast
    f = rnd(2, 3, 8, 8, seed=6)
    assert torch.equal(ops.cast(f, torch.bfloat16), f.to(torch.bfloat16).contiguous(memory_format=CL))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("ncls", [2, 4])
def test_head(dtype, ncls):
    ops = _ops()
    x = rnd(2, 32, 16, 16, dtype=dtype, seed=1).contiguous(memory_format=CL).requires_grad_(True)
    w = rnd(ncls, 32, 1, 1, seed=2, scale=0.2).requires_grad_(True)
    b = rnd(ncls, seed=3).requires_grad_(True)
    y = ops.head(x, w, b)
    assert y.dtype == torch.float32
    gy = rnd(*y.shape, seed=4)
    y.backward(gy)
    xr, wr, br = x.detach().float().requires_grad_(True), w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, br)
    yr.backward(gy)
    t = tol(dtype)
    assert rel(y, yr) < 1e-4 and rel(x.grad, xr.grad) < t and rel(w.grad, wr.grad) < 1e-3 and rel(b.grad, br.grad) < 1e-4


@pytest.mark.parametrize("loss_str,ncls,post", [("focal+dice", 2, False), ("dice", 2, False), ("focal", 4, True),
                                                 ("focal+dice", 4, True), ("ce+ohem", 2, False), ("ce", 4, True)])
@pytest.mark.parametrize("lstride", [1, 2])
def test_seg_loss_vs_oracle(loss_str, ncls, post, lstride):
    from oracle import functional as OF
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    logits = (torch.randn(2, ncls, 16, 24, generator=g) * 2).cuda().requires_grad_(True)
    hi = 5 if post else 2
    labels = torch.randint(0, hi, (2, 16 * lstride, 24 * lstride), generator=g, dtype=torch.uint8).cuda()
    loss = ops.seg_loss(logits, labels, loss_str, post, weight=0.5, lstride=lstride)
    loss.backward()
    lr = logits.detach().cpu().clone().requires_grad_(True)
    ref = 0.5 * OF.loss_forward(lr, labels.cpu()[:, ::lstride, ::lstride], loss_str, post)
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    assert rel(logits.grad, lr.grad) < 1e-4


def test_f1_and_postprocess_bit_exact():
    from oracle import functional as OF
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    # localisation metric (2 classes)
    logits = torch.randn(2, 2, 32, 32, generator=g)
    logits[0, :, 0, :4] = 1.25  # exact ties -> class 0
    labels = torch.randint(0, 2, (2, 32, 32), generator=g, dtype=torch.uint8)
    ctr = torch.zeros(3, dtype=torch.int64, device="cuda")
    pred = torch.empty((2, 32, 32), dtype=torch.uint8, device="cuda")
    ops.f1_update(logits.cuda(), labels.cuda(), 2, ctr, pred)
    tp, fp, fn = OF.f1_counters(logits, labels, 2)
    assert ctr.cpu().tolist() == [int(tp[0]), int(fp[0]), int(fn[0])]
    assert torch.equal(pred.cpu(), torch.argmax(logits, 1).to(torch.uint8))
    # damage metric (5 classes, 4 logits, building pixels only)
    logits4 = torch.randn(2, 4, 32, 32, generator=g)
    labels5 = torch.randint(0, 5, (2, 32, 32), generator=g, dtype=torch.uint8)
    ctr4 = torch.zeros(12, dtype=torch.int64, device="cuda")
    ops.f1_update(logits4.cuda(), labels5.cuda(), 5, ctr4)
    tp, fp, fn = OF.f1_counters(logits4, labels5, 5)
    assert ctr4.cpu().tolist() == [*map(int, tp), *map(int, fp), *map(int, fn)]
    # post-process from probabilities: bit-exact against the numpy rule
    loc = torch.rand(64, 64, generator=g)
    dmg = torch.softmax(torch.randn(4, 64, 64, generator=g), 0)
    dmg[:, 0, :8] = 0.25
    pre, post = ops.post_process_probs(loc.cuda(), dmg.cuda())
    pre_r, post_r = OF.post_process(loc.numpy(), dmg.numpy())
    assert np.array_equal(pre.cpu().numpy(), pre_r) and np.array_equal(post.cpu().numpy(), post_r)
    # fused from logits: identical away from the two thresholds
    loc_l = torch.randn(1, 2, 64, 64, generator=g) * 3
    dmg_l = torch.randn(1, 4, 64, 64, generator=g)
    pre2, post2 = ops.post_process(loc_l.cuda(), dmg_l.cuda())
    prob = torch.sigmoid(loc_l[0, 1])
    safe = ((prob - 0.3).abs() > 1e-5) & ((prob - 0.1).abs() > 1e-5)
    pre_r2, post_r2 = OF.post_process(prob.numpy(), torch.softmax(dmg_l[0], 0).numpy())
    assert np.array_equal(pre2[0].cpu().numpy()[safe.numpy()], pre_r2[safe.numpy()])
    assert np.array_equal(post2[0].cpu().numpy()[safe.numpy()], post_r2[safe.numpy()])


def test_mean4_and_normalize_and_adamw():
    from oracle import functional as OF
    ops = _ops()
    ts = [rnd(2, 2, 8, 8, seed=i) for i in range(4)]
    m = ops.mean4(*ts)
    ref = ts[0].clone()
    for t in ts[1:]:
        ref += t
    ref /= 4
    assert torch.equal(m, ref.contiguous(memory_format=CL))
    g = torch.Generator().manual_seed(1)
    pre = torch.randint(0, 256, (2, 16, 16, 3), generator=g, dtype=torch.uint8)
    post = torch.randint(0, 256, (2, 16, 16, 3), generator=g, dtype=torch.uint8)
    out = ops.normalize_tiles(pre.cuda(), post.cuda(), torch.float32)
    ref = np.stack([np.concatenate([OF.normalize_tile(pre[i].numpy()), OF.normalize_tile(post[i].numpy())], 0) for i in range(2)])
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-6
    out3 = ops.normalize_tiles(pre.cuda(), None, torch.bfloat16)
    assert out3.shape == (2, 3, 16, 16) and out3.dtype == torch.bfloat16
    # AdamW against torch.optim.AdamW
    p = rnd(1000, seed=1)
    pr = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([pr], lr=3e-4, weight_decay=0.01)
    m_, v_ = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        gq = rnd(1000, seed=10 + step)
        pr.grad = gq.clone()
        opt.step()
        ops.adamw_step(p, gq, m_, v_, 3e-4, 0.9, 0.999, 1e-8, 0.01, step)
    assert rel(p, pr.data) < 1e-5


@pytest.mark.parametrize("ncls", [2, 4])
def test_save_probs_values_vs_reference_formula(ncls):
    """Model.save (plt.py:126-131): sigmoid(pred[:, 1]) for the localisation head, softmax(pred, 1) (planar NCHW, the layout
    np.save receives) for the damage head -- VALUES against torch, not just range / shape."""
    ops = _ops()
    logits = rnd(3, ncls, 40, 56, seed=21, scale=3.0).contiguous(memory_format=CL)
    got = ops.save_probs(logits)
    ref = torch.sigmoid(logits[:, 1]) if ncls == 2 else torch.softmax(logits, 1)
    assert got.shape == ref.shape and got.dtype == torch.float32 and got.is_contiguous()
    assert float((got - ref).abs().max()) < 2e-6
    if ncls == 4:
        assert float((got.sum(1) - 1).abs().max()) < 1e-5
    # bf16 logits (what the tensor-core path hands over) go through the same kernel after an exact up-cast
    got16 = ops.save_probs(logits.to(torch.bfloat16))
    ref16 = torch.sigmoid(logits.to(torch.bfloat16).float()[:, 1]) if ncls == 2 else torch.softmax(logits.to(torch.bfloat16).float(), 1)
    assert float((got16 - ref16).abs().max()) < 2e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("bins,hw", [(1, (5, 7)), (2, (4, 4)), (3, (8, 10)), (6, (16, 16)), (6, (7, 9))])
def test_adaptive_avg_pool(dtype, bins, hw):
    """PPM's nn.AdaptiveAvgPool2d (layers.py:13) forward / backward vs torch."""
    ops = _ops()
    x = rnd(2, 24, *hw, dtype=dtype, seed=3).contiguous(memory_format=CL).requires_grad_(True)
    y = ops.adaptive_avg_pool2d(x, bins)
    gy = rnd(*y.shape, dtype=dtype, seed=4).contiguous(memory_format=CL)
    y.backward(gy)
    xr = x.detach().float().requires_grad_(True)
    yr = F.adaptive_avg_pool2d(xr, bins)
    yr.backward(gy.float())
    assert rel(y, yr) < tol(dtype) and rel(x.grad, xr.grad) < tol(dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("hw,out", [((1, 1), (8, 8)), ((2, 2), (8, 8)), ((3, 6), (16, 16)), ((16, 16), (32, 32)), ((8, 8), (64, 64)),
                                    ((16, 12), (5, 7))])
def test_bilinear_align_corners(dtype, hw, out):
    """F.interpolate(mode='bilinear', align_corners=True) (layers.py:27,154,188) forward / backward vs torch."""
    ops = _ops()
    x = rnd(2, 12, *hw, dtype=dtype, seed=7).contiguous(memory_format=CL).requires_grad_(True)
    y = ops.bilinear(x, out)
    gy = rnd(*y.shape, dtype=dtype, seed=8).contiguous(memory_format=CL)
    y.backward(gy)
    xr = x.detach().float().requires_grad_(True)
    yr = F.interpolate(xr, out, mode="bilinear", align_corners=True)
    yr.backward(gy.float())
    assert rel(y, yr) < tol(dtype) and rel(x.grad, xr.grad) < tol(dtype)


@pytest.mark.parametrize("loss_str,post", [("mse", True), ("coral", True), ("coral", False), ("mse", False)])
def test_ordinal_heads_vs_oracle(loss_str, post):
    """'mse' / 'coral' damage heads: loss + gradient (loss.py:54-65,86-94) and label decoding + F1 counters (utils/f1.py:7-42)."""
    from oracle import functional as OF
    ops = _ops()
    nl = 1 if loss_str == "mse" else 3
    g = torch.Generator().manual_seed(13)
    logits = (torch.randn(3, nl, 24, 40, generator=g) * 2).cuda().contiguous(memory_format=CL).requires_grad_(True)
    hi = 5 if post else 4
    labels = torch.randint(0, hi, (3, 24, 40), generator=g, dtype=torch.uint8).cuda()
    loss = ops.seg_loss(logits, labels, loss_str, post, weight=0.5)
    loss.backward()
    lr = logits.detach().cpu().requires_grad_(True)
    ref = 0.5 * OF.loss_forward(lr, labels.cpu(), loss_str, post)
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    assert rel(logits.grad, lr.grad) < 1e-5
    # decoding: exact integers
    got = ops.ordinal_labels(logits, loss_str, want_u8=True).cpu().long()
    want = OF.convert_to_labels(loss_str, logits.detach().cpu()).long()
    assert torch.equal(got, want)
    if post:
        counters = torch.zeros(12, dtype=torch.int64, device="cuda")
        ops.ordinal_labels(logits, loss_str, labels, counters)
        t = labels.cpu().long()
        keep = t > 0
        for c in range(1, 5):
            tp = int(((want == c) & (t == c) & keep).sum())
            fp = int(((want == c) & (t != c) & keep).sum())
            fn = int(((want != c) & (t == c) & keep).sum())
            assert (int(counters[c - 1]), int(counters[4 + c - 1]), int(counters[8 + c - 1])) == (tp, fp, fn)


def test_cat_unet_matches_oracle():
    """CatUNet (unet.py:554-560): the reference's constructor raises; the 6-channel stem it intended is checked
    against the oracle's functional U-Net given the same weights."""
    import argparse

    from oracle import functional as OF
    from xview2_b200.model.unet import get_dmg_unet
    ns = argparse.Namespace(ppm=False, aspp=False, dilation=1, no_skip=False, interpolate=False, attention=False, dec_interp=False,
                            deep_supervision=False, loss_str="focal+dice", encoder="resnest50", dmg_model="cat", type="post",
                            tta=False, precision=32)
    model = get_dmg_unet(ns)
    shapes = {k: (tuple(v.shape), v.dtype) for k, v in model.state_dict().items()}
    state = OF.deterministic_state(shapes, 1)
    model.load_state_dict(state, strict=True)
    model = model.cuda().train()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(4, 6, 64, 64, generator=g)
    out = model(x.cuda())
    ref = OF.model_forward({k: v.clone() for k, v in state.items()}, x, True, ns)
    assert out.shape == ref.shape == (4, 4, 64, 64)
    assert rel(out, ref) < 1e-3


def test_batched_repack_equals_single_pack():
    """xv2_pack_weights_batched (one launch per step for every packed copy of the model) writes exactly what xv2_pack_weight
    (one launch per weight) writes, for every mode / dtype / group count the networks use."""
    ops = _ops()
    cases = [((128, 64, 3, 3), 0, torch.bfloat16, 1), ((128, 64, 3, 3), 1, torch.bfloat16, 1), ((256, 64, 3, 3), 1, torch.bfloat16, 2),
             ((64, 256, 1, 1), 1, torch.bfloat16, 1), ((128, 32, 2, 2), 2, torch.bfloat16, 1), ((512, 64, 1, 1), 1, torch.float32, 1),
             ((32, 3, 3, 3), 0, torch.bfloat16, 1), ((48, 40, 3, 3), 1, torch.bfloat16, 1), ((2, 32, 1, 1), 1, torch.bfloat16, 1)]
    ops.clear_weight_cache()
    weights, single = [], []
    for i, (shape, mode, dtype, groups) in enumerate(cases):
        w = rnd(*shape, seed=20 + i).contiguous(memory_format=CL)
        weights.append(w)
        single.append(ops.pack_weight(w, mode, dtype, groups).clone())  # element-wise kernel
    for w in weights:
        w.mul_(1.0)  # bump the version: the cached copies are stale now
    for (shape, mode, dtype, groups), w in zip(cases, weights):
        ops.pack_weight(w, mode, dtype, groups).zero_()  # re-packs lazily (element-wise), then wiped so the batched launch must write
    ops.clear_weight_cache()
    ops.repack_all()
    torch.cuda.synchronize()
    for (shape, mode, dtype, groups), w, ref in zip(cases, weights, single):
        got = ops.pack_weight(w, mode, dtype, groups)  # valid stamp: returned as the batched launch left it
        assert torch.equal(got, ref), (shape, mode, dtype, groups)


def test_folded_split_attention_entries_equal_their_unfolded_pieces():
    """The per-bottleneck launches folded away this round, checked at the C ABI: xv2_splat_bn_gap_fin == xv2_bn_finalize +
    xv2_splat_bn_gap (coefficients / running statistics bit-exact, pooled vector to fp32 rounding: the folded entry accumulates
    in fp64), xv2_splat_fc_bwd_fused == xv2_splat_bn_bwd_datt + xv2_splat_fc_bwd + xv2_splat_bn_bwd_red (bit-exact), and its
    accumulate mode adds onto what the destinations already hold."""
    from xview2_b200 import lib
    from xview2_b200.lib import BF16, call, ptr
    lib.init(torch.cuda.current_device())
    n, c, hw, inter = 4, 64, 32 * 32, 32
    c2 = 2 * c
    f32 = dict(dtype=torch.float32, device="cuda")
    z = rnd(n, hw, c2, dtype=torch.bfloat16, seed=1)
    dout = rnd(n, hw, c, dtype=torch.bfloat16, seed=2)
    g0, b0 = rnd(c2, seed=3) * 0.2 + 1, rnd(c2, seed=4) * 0.2
    stats = torch.zeros(2 * c2, dtype=torch.float64, device="cuda")
    call("xv2_bn_stats", ptr(z), n * hw, c2, BF16, ptr(stats))
    # ---- forward: finalize + gap vs the folded entry
    rm_a, rv_a = torch.zeros(c2, **f32), torch.ones(c2, **f32)
    rm_b, rv_b = rm_a.clone(), rv_a.clone()
    coef_a, coef_b = torch.empty(4, c2, **f32), torch.empty(4, c2, **f32)
    call("xv2_bn_finalize", ptr(stats), n * hw, c2, ptr(g0), ptr(b0), ptr(rm_a), ptr(rv_a), 0.1, 1e-5, ptr(coef_a[0]), ptr(coef_a[1]),
         ptr(coef_a[2]), ptr(coef_a[3]))
    gap_a = torch.empty(n, c, **f32)
    call("xv2_splat_bn_gap", ptr(z), ptr(coef_a[2]), ptr(coef_a[3]), ptr(gap_a), n, hw, c)
    gap_b, gap_acc = torch.empty(n, c, **f32), torch.zeros(n, c, dtype=torch.float64, device="cuda")
    call("xv2_splat_bn_gap_fin", ptr(z), ptr(stats), n * hw, ptr(g0), ptr(b0), ptr(rm_b), ptr(rv_b), 0.1, 1e-5, ptr(coef_b), ptr(gap_b),
         ptr(gap_acc), n, hw, c)
    assert torch.equal(coef_a, coef_b) and torch.equal(rm_a, rm_b) and torch.equal(rv_a, rv_b)
    assert rel(gap_b, gap_a) < 1e-5
    # ---- the FC chain forward (shared by both paths) for att / a1 / z1 / coef
    w1, b1 = rnd(inter, c, seed=5) * 0.1, rnd(inter, seed=6) * 0.1
    w2, b2 = rnd(c2, inter, seed=7) * 0.1, rnd(c2, seed=8) * 0.1
    g1, be1 = rnd(inter, seed=9) * 0.2 + 1, rnd(inter, seed=10) * 0.2
    rm1, rv1 = torch.zeros(inter, **f32), torch.ones(inter, **f32)
    z1, a1, coef1, att = torch.empty(n, inter, **f32), torch.empty(n, inter, **f32), torch.empty(4, inter, **f32), torch.empty(n, c2, **f32)
    call("xv2_splat_fc_fwd", ptr(gap_b), ptr(w1), ptr(b1), ptr(g1), ptr(be1), ptr(rm1), ptr(rv1), 0.1, 1e-5, 1, ptr(w2), ptr(b2), ptr(z1),
         ptr(a1), ptr(coef1), ptr(att), n, c, inter)
    part = torch.zeros(4, n, c2, dtype=torch.float64, device="cuda")
    call("xv2_splat_bn_bwd_partials", ptr(z), ptr(dout), ptr(coef_a[2]), ptr(coef_a[3]), ptr(part), n, hw, c)
    w2t, w1t = w2.t().contiguous(), w1.t().contiguous()

    def outs(fill=0.0):
        o = {"dz2": torch.empty(n, c2, **f32), "dz1": torch.empty(n, inter, **f32), "dw2": torch.full((c2, inter), fill, **f32),
             "db2": torch.full((c2,), fill, **f32), "dw1": torch.full((inter, c), fill, **f32), "db1": torch.full((inter,), fill, **f32),
             "dgamma": torch.full((inter,), fill, **f32), "dbeta": torch.full((inter,), fill, **f32), "dgap": torch.empty(n, c, **f32),
             "red": torch.empty(2 * c2, dtype=torch.float64, device="cuda")}
        return o

    # ---- backward: datt + fc_bwd + red vs the folded entry
    a = outs()
    datt = torch.empty(n, c2, **f32)
    call("xv2_splat_bn_bwd_datt", ptr(part), ptr(coef_a[2]), ptr(coef_a[3]), ptr(datt), n, c)
    call("xv2_splat_fc_bwd", ptr(att), ptr(datt), ptr(a1), ptr(z1), ptr(coef1), ptr(g1), ptr(gap_b), ptr(w2t), ptr(w1t), 1, ptr(a["dz2"]),
         ptr(a["dz1"]), ptr(a["dw2"]), ptr(a["db2"]), ptr(a["dw1"]), ptr(a["db1"]), ptr(a["dgamma"]), ptr(a["dbeta"]), ptr(a["dgap"]), n, c,
         inter)
    call("xv2_splat_bn_bwd_red", ptr(part), ptr(att), ptr(a["dgap"]), ptr(coef_a[0]), ptr(coef_a[1]), ptr(a["red"]), n, hw, c)

    def fused(o, accumulate):
        call("xv2_splat_fc_bwd_fused", ptr(att), ptr(part), ptr(coef_a[2]), ptr(coef_a[3]), ptr(coef_a[0]), ptr(coef_a[1]), hw, ptr(a1),
             ptr(z1), ptr(coef1), ptr(g1), ptr(gap_b), ptr(w2t), ptr(w1t), 1, ptr(o["dz2"]), ptr(o["dz1"]), ptr(o["dw2"]), ptr(o["db2"]),
             ptr(o["dw1"]), ptr(o["db1"]), ptr(o["dgamma"]), ptr(o["dbeta"]), ptr(o["dgap"]), ptr(o["red"]), accumulate, n, c, inter)

    b = outs(7.0)  # garbage in the destinations: accumulate = 0 must overwrite it
    fused(b, 0)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    acc = outs(0.5)
    fused(acc, 1)
    for k in ("dw2", "db2", "dw1", "db1", "dgamma", "dbeta"):
        assert torch.equal(acc[k], a[k] + 0.5), k
    assert torch.equal(acc["dgap"], a["dgap"]) and torch.equal(acc["red"], a["red"])
