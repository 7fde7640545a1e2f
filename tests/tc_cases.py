"""Tensor-core (tcgen05) convolution cases, shared by tests/test_tc_gpu.py and tools/tc_probe.py.

Every case runs the op through xview2_b200.ops with the tensor-core path enabled and compares forward, data gradient
and weight gradient with torch's fp32 convolution on the same bf16-rounded inputs.  `lib.launches()` bookkeeping plus a
direct check of the entry-point return code make sure the tcgen05 kernel (not the SIMT fallback) produced the result.
"""
import torch
import torch.nn.functional as F

CL = torch.channels_last

# name: (kind, n, c0, c1, h, w, k, r, dil, groups)
CASES = {
    "c64_k64_3x3_w16": ("conv", 2, 64, 0, 16, 16, 64, 3, 1, 1),
    "c64_k256_1x1_w32": ("conv", 2, 64, 0, 32, 32, 256, 1, 1, 1),
    "c32_k32_3x3_w128": ("conv", 1, 32, 0, 8, 128, 32, 3, 1, 1),
    "c32_k64_3x3_w64": ("conv", 2, 32, 0, 8, 64, 64, 3, 1, 1),
    "c128_k512_1x1_w8": ("conv", 2, 128, 0, 16, 8, 512, 1, 1, 1),
    "g2_c64_k128_3x3": ("conv", 2, 64, 0, 16, 16, 128, 3, 1, 2),       # SplAt radix conv, 32 ch / group
    "g2_c128_k256_3x3": ("conv", 2, 128, 0, 16, 16, 256, 3, 1, 2),
    "cat_64_128_k64_3x3": ("conv", 2, 64, 128, 16, 16, 64, 3, 1, 1),   # decoder conv over (up, skip)
    "cat_32_32_k32_3x3": ("conv", 2, 32, 32, 16, 32, 32, 3, 1, 1),
    "dil2_c64_k64": ("conv", 1, 64, 0, 16, 16, 64, 3, 2, 1),
    "c256_k96_3x3_w256": ("conv", 1, 256, 0, 2, 256, 96, 3, 1, 1),     # N tile not a power of two
    # row-strip kernel shapes (w % 128 == 0, <= 128 channels per group): ring wrap, pieces crossing columns / images
    "strip_c32_k32_w256": ("conv", 3, 32, 0, 100, 256, 32, 3, 1, 1),
    "strip_c64_k64_w128": ("conv", 2, 64, 0, 37, 128, 64, 3, 1, 1),
    "strip_c32_k64_w128": ("conv", 2, 32, 0, 64, 128, 64, 3, 1, 1),
    "strip_cat_64_64_k64_w128": ("conv", 2, 64, 64, 48, 128, 64, 3, 1, 1),
    "strip_g2_c64_k128_w128": ("conv", 2, 64, 0, 40, 128, 128, 3, 1, 2),
    "strip_g2_c128_k256_w128": ("conv", 1, 128, 0, 33, 128, 256, 3, 1, 2),
    # strip weight-gradient shapes: narrow images, k blocks of 128, two sources with 32-channel chunks
    "wg_c96_k256_w32": ("conv", 2, 96, 0, 20, 32, 256, 3, 1, 1),
    "wg_cat_32_64_k128_w64": ("conv", 2, 32, 64, 24, 64, 128, 3, 1, 1),
    "wg_g2_c64_k256_w16": ("conv", 3, 64, 0, 16, 16, 256, 3, 1, 2),
    "convt_c128_k64": ("convt", 2, 128, 0, 16, 16, 64, 2, 1, 1),
    "convt_c64_k32_w64": ("convt", 1, 64, 0, 8, 64, 32, 2, 1, 1),
    "convt_c2048_k512": ("convt", 1, 2048, 0, 16, 8, 512, 2, 1, 1),
    "convt_c64_k32_w256": ("convt", 2, 64, 0, 6, 256, 32, 2, 1, 1),    # all four taps in one N tile (dec_l5 shape), row tiles
}


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def ulp_excess(out_bf16, ref_fp32):
    """How far a bf16 result is from the fp32 reference BEYOND one bf16 ulp of each element (2^-8 relative: rounding of the
    fp32 accumulator plus a one-ulp difference in where the accumulation-order noise lands), relative to the largest
    reference magnitude.  ~1e-5 for a correct kernel (fp32 accumulation-order noise); a dropped K block shows up as O(0.1)."""
    a, b = out_bf16.detach().double().cpu(), ref_fp32.detach().double().cpu()
    exc = ((a - b).abs() - b.abs() * 2.0 ** -8).clamp_min(0.0)
    return float(exc.max() / b.abs().max().clamp_min(1e-12))


def _rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def run_case(name, check_wgrad=True, check_dgrad=True):
    """Returns dict(fwd=, dgrad=, wgrad=) of relative errors; raises if the tensor-core entry declined the shape."""
    from xview2_b200 import lib, ops
    kind, n, c0, c1, h, w, k, r, dil, groups = CASES[name]
    ops.USE_TENSOR_CORES = True
    calls = []
    orig_call = ops.call

    def spy(fn, *a, **kw):
        rc = orig_call(fn, *a, **kw)
        calls.append((fn, rc))
        return rc

    ops.call = spy
    try:
        x = _rnd((n, c0, h, w), 1).to(torch.bfloat16).contiguous(memory_format=CL).requires_grad_(True)
        x2 = _rnd((n, c1, h, w), 2).to(torch.bfloat16).contiguous(memory_format=CL).requires_grad_(True) if c1 else None
        if kind == "conv":
            cg = (c0 + c1) // groups
            wt = _rnd((k, cg, r, r), 3, (2.0 / (cg * r * r)) ** 0.5).contiguous(memory_format=CL).requires_grad_(True)
            pad = dil * (r - 1) // 2
            y = ops.conv2d(x, wt, None, 1, pad, dil, groups, x2)
        else:
            wt = _rnd((c0, k, 2, 2), 3, (1.0 / c0) ** 0.5).contiguous(memory_format=CL).requires_grad_(True)
            y = ops.conv_transpose2x2(x, wt)
        gy = _rnd(tuple(y.shape), 4).to(torch.bfloat16).contiguous(memory_format=CL)
        y.backward(gy)
        torch.cuda.synchronize()
    finally:
        ops.call = orig_call
    tc_calls = [(f, rc) for f, rc in calls if f in ("xv2_conv_tc", "xv2_wgrad_tc")]
    declined = [(f, rc) for f, rc in tc_calls if rc != 0]
    if declined or not tc_calls:
        raise AssertionError(f"{name}: tensor-core entry declined or not used: {tc_calls} ({lib.last_error()})")
    xr = x.detach().float().requires_grad_(True)
    wr = wt.detach().to(torch.bfloat16).float().requires_grad_(True)
    if kind == "conv":
        src = xr
        x2r = None
        if c1:
            x2r = x2.detach().float().requires_grad_(True)
            src = torch.cat((xr, x2r), 1)
        yr = F.conv2d(src, wr, None, 1, pad, dil, groups)
    else:
        x2r = None
        yr = F.conv_transpose2d(xr, wr, None, 2)
    yr.backward(gy.float())
    out = {"fwd": rel(y, yr), "fwd_ulp": ulp_excess(y, yr)}
    if check_dgrad:
        out["dgrad"] = rel(x.grad, xr.grad)
        out["dgrad_ulp"] = ulp_excess(x.grad, xr.grad)
        if c1:
            out["dgrad2"] = rel(x2.grad, x2r.grad)
            out["dgrad2_ulp"] = ulp_excess(x2.grad, x2r.grad)
    if check_wgrad:
        out["wgrad"] = rel(wt.grad, wr.grad)
    return out


def run_f32_case(name):
    """The same tcgen05 main loop with its fp32-output epilogue, called straight through the C ABI (xv2_conv_tc with
    out_dtype = XV2_F32): exact bf16 operands, fp32 TMEM accumulation, no output rounding -> the only difference from
    torch's fp32 convolution on the same operands is the summation order.  Returns the relative error (max |d| / max |ref|)."""
    from xview2_b200 import lib, ops
    from xview2_b200.lib import BF16, F32, TcConv, call, ptr
    kind, n, c0, c1, h, w, k, r, dil, groups = CASES[name]
    lib.init(torch.cuda.current_device())
    x = _rnd((n, c0, h, w), 1).to(torch.bfloat16).contiguous(memory_format=CL)
    x2 = _rnd((n, c1, h, w), 2).to(torch.bfloat16).contiguous(memory_format=CL) if c1 else None
    if kind == "conv":
        cg = (c0 + c1) // groups
        wt = _rnd((k, cg, r, r), 3, (2.0 / (cg * r * r)) ** 0.5).contiguous(memory_format=CL)
        pad = dil * (r - 1) // 2
        wp = ops.pack_weight(wt, 0, torch.bfloat16, groups)
        out = torch.empty((n, k, h, w), dtype=torch.float32, device="cuda").contiguous(memory_format=CL)
        p = TcConv(n, h, w, c0, c1, 0, 0, k, r, r, pad, dil, groups, 0, F32, 0)
        call("xv2_conv_tc", p, ptr(x), ptr(x2), ptr(wp), None, ptr(out), None)
        src = x.float() if x2 is None else torch.cat((x.float(), x2.float()), 1)
        ref = F.conv2d(src, wt.to(torch.bfloat16).float(), None, 1, pad, dil, groups)
    else:
        wt = _rnd((c0, k, 2, 2), 3, (1.0 / c0) ** 0.5).contiguous(memory_format=CL)
        wp = ops.pack_weight(wt, 2, torch.bfloat16)
        out = torch.empty((n, k, 2 * h, 2 * w), dtype=torch.float32, device="cuda").contiguous(memory_format=CL)
        p = TcConv(n, h, w, c0, 0, 0, 0, k, 1, 1, 0, 1, 1, 1, F32, 0)
        call("xv2_conv_tc", p, ptr(x), None, ptr(wp), None, ptr(out), None)
        ref = F.conv_transpose2d(x.float(), wt.to(torch.bfloat16).float(), None, 2)
    torch.cuda.synchronize()
    return rel(out, ref)
