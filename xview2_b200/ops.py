"""Differentiable operators of the segmentation hot path, each a thin torch.autograd.Function over libxv2 entry points.

Tensors are logical NCHW with channels-last strides (physical NHWC), bf16 (tensor-core path) or fp32 (parity path).
PyTorch supplies device memory, streams and the autograd tape; every arithmetic kernel is ours (xview2_b200/csrc).
"""
import torch

from . import lib
from .lib import ACT_LRELU, ACT_NONE, ACT_RELU, BF16, F32, ConvGeom, TcConv, call, dtype_code, ptr

CL = torch.channels_last

# Set to False to force the SIMT path everywhere (used by tests to cross-check the tensor-core kernels).
USE_TENSOR_CORES = True
# Fused streaming paths (A/B switches for tests and profiling; XV2_NO_FUSE=1 turns both off):
#   FUSE_SPLAT_BN: bn0 + ReLU folded into the split-attention kernels (csrc/splat_fused.cu)
#   FUSE_TAIL    : last decoder BatchNorm + LeakyReLU folded into the 1x1 head (csrc/fused_tail.cu)
import os as _os0
FUSE_SPLAT_BN = _os0.environ.get("XV2_NO_FUSE", "0") != "1"
FUSE_TAIL = _os0.environ.get("XV2_NO_FUSE", "0") != "1"


def nhwc(t):
    """Returns `t` with channels-last strides (no copy if it already has them)."""
    if t.dim() != 4:
        raise lib.Xv2Error(f"expected a 4-D activation, got shape {tuple(t.shape)}")
    if t.is_contiguous(memory_format=CL):
        # size-1 dims make the check ambiguous; normalise strides so data_ptr arithmetic is NHWC for sure
        n, c, h, w = t.shape
        if t.stride() == (h * w * c, 1, w * c, c):
            return t
        return t.as_strided((n, c, h, w), (h * w * c, 1, w * c, c))
    return t.contiguous(memory_format=CL)


def empty_act(n, c, h, w, dtype, device):
    return torch.empty((n, c, h, w), dtype=dtype, device=device, memory_format=CL)


def _require_cuda(t):
    if not t.is_cuda:
        raise lib.Xv2Error("xview2_b200 operators run on CUDA tensors only (no CPU fallback)")
    lib.init(t.device.index)


# ---------------------------------------------------------------------------------------------------------------
# packed weights (bf16 / fp32 copies in kernel order), cached per parameter version
# ---------------------------------------------------------------------------------------------------------------
_pack_epoch = 0  # bumped whenever parameters are rewritten behind autograd's back (optimizer kernels, broadcasts)


def _weight_phys(weight):
    """fp32 parameter in its physical channels-last order [A][R][S][B]; converts once if needed."""
    if weight.dtype != torch.float32:
        raise lib.Xv2Error("master weights must be fp32")
    return nhwc(weight)


def pack_weight(weight, mode, dtype, groups=1):
    """mode 0: [K][R][S][Cg]; mode 1: dgrad order; mode 2: transposed-conv GEMM rows.

    The packed copy lives ON the parameter object (so it dies with it -- a cache keyed by data_ptr would hand a new
    tensor that reuses the address a stale copy) and is valid for one (tensor version, pack epoch, storage address)."""
    cache = weight.__dict__.get("_xv2_pack")
    if cache is None:
        cache = {}
        weight.__dict__["_xv2_pack"] = cache
    stamp = (weight._version, _pack_epoch, weight.data_ptr())
    hit = cache.get((mode, dtype, groups))
    if hit is not None and hit[0] == stamp and hit[1].device == weight.device:
        return hit[1]
    w = _weight_phys(weight)
    a, b, r, s = w.shape
    out = hit[1] if hit is not None and hit[1].device == weight.device and hit[1].numel() == w.numel() else \
        torch.empty(w.numel(), dtype=dtype, device=w.device)
    call("xv2_pack_weight", ptr(w), ptr(out), a, r, s, b, groups, mode, dtype_code(out))
    cache[(mode, dtype, groups)] = (stamp, out)
    import weakref
    _pack_registry[(id(weight), mode, dtype, groups)] = (weakref.ref(weight), out)
    return out


def clear_weight_cache():
    """Invalidates every packed weight copy (called after kernels rewrite the master weights in place)."""
    global _pack_epoch
    _pack_epoch += 1


_pack_registry = {}   # (id(weight), mode, dtype, groups) -> (weakref to weight, packed tensor)
_pack_table = None    # (registry size, device job table, keep-alive list)


def repack_all():
    """Refreshes EVERY packed copy known so far with one launch (xv2_pack_weights_batched) and stamps them valid for the
    current parameter values.  Called by the fused optimizers right after they rewrote the flat master buffer, so the
    per-layer lazy re-pack (171 launches per step at ResNeSt-50) disappears from the step."""
    global _pack_table
    import numpy as np

    live = []
    for key, (ref, out) in list(_pack_registry.items()):
        w = ref()
        if w is None or not w.is_cuda or w.__dict__.get("_xv2_pack", {}).get(key[1:], (None, None))[1] is not out:
            del _pack_registry[key]
            continue
        live.append((key, w, out))
    if not live:
        return
    sig = tuple((k, w.data_ptr(), o.data_ptr()) for k, w, o in live)
    if _pack_table is None or _pack_table[0] != sig:
        rows = np.zeros((len(live), 6), dtype=np.int64)  # sizeof(xv2_pack_job) = 8 + 8 + 8 * 4 = 48 bytes
        for i, (key, w, out) in enumerate(live):
            _, mode, dtype, groups = key
            phys = _weight_phys(w)
            a, b, r, s = phys.shape
            rows[i, 0] = phys.data_ptr()
            rows[i, 1] = out.data_ptr()
            ints = np.array([a, r, s, b, groups, mode, dtype_code(out), 0], dtype=np.int32)
            rows[i, 2:6] = ints.view(np.int64)
        table = torch.from_numpy(rows).to(live[0][1].device)
        _pack_table = (sig, table)
    call("xv2_pack_weights_batched", ptr(_pack_table[1]), len(live))
    for key, w, out in live:
        w.__dict__["_xv2_pack"][key[1:]] = ((w._version, _pack_epoch, w.data_ptr()), out)


# ---------------------------------------------------------------------------------------------------------------
# convolution
# ---------------------------------------------------------------------------------------------------------------
def _grad_buffer(weight):
    """Where a weight-gradient kernel may ACCUMULATE directly: the parameter's own .grad when it is the zero-initialised fp32
    view FlatParams installed (physical [K][R][S][C], same order the kernels write).  Returns (buffer, direct): with
    direct=True the autograd function returns None for this weight, so autograd launches no zeros / add kernels for it."""
    g = weight.grad
    if (g is not None and g.dtype == torch.float32 and g.is_cuda and g.shape == weight.shape and
            g.is_contiguous(memory_format=CL) and getattr(weight, "_xv2_flat", False)):
        return g, True
    return torch.zeros(weight.shape, dtype=torch.float32, device=weight.device).contiguous(memory_format=CL), False


# Weight gradients are leaves of the backward pass (only the optimizer reads them), so they are issued on a SIDE stream and
# overlap the batch-norm / data-gradient chain of the layers below (tensor-pipe work next to HBM-bound work).
import os as _os

# Opt-in (bench.py / Trainer.fit switch it on): whoever reads .grad must call sync_side_streams() first -- FlatParams'
# all_reduce_grads(), zero_grad() and the fused optimizers' step() do.
WGRAD_SIDE_STREAM = False


def enable_wgrad_side_stream(on=True):
    global WGRAD_SIDE_STREAM
    # measured (r01, C2 step): no gain while the wgrad and BN kernels cannot co-reside on an SM (185 KB + 112 KB of shared
    # memory) and record_stream() churns the allocator -> off unless XV2_WGRAD_STREAM=1
    WGRAD_SIDE_STREAM = bool(on) and _os.environ.get("XV2_WGRAD_STREAM", "0") == "1"
_side_streams = {}
_side_dirty = set()


def _side_stream(device):
    st = _side_streams.get(device.index)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _side_streams[device.index] = st
    return st


class _OnSide:
    """Context: kernels launched inside run on the side stream after everything already queued on the current stream;
    the listed tensors are kept alive (allocator-wise) until the side stream has consumed them."""

    def __init__(self, device, *tensors):
        self.side = _side_stream(device)
        self.tensors = [t for t in tensors if t is not None]
        self.ctx = None

    def __enter__(self):
        self.side.wait_stream(torch.cuda.current_stream())
        for t in self.tensors:
            t.record_stream(self.side)
        _side_dirty.add(self.side)
        self.ctx = torch.cuda.stream(self.side)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        return self.ctx.__exit__(*exc)


def sync_side_streams():
    """The current stream waits for every weight gradient issued on a side stream (called before the gradient all-reduce /
    optimizer step / anything that reads .grad)."""
    cur = torch.cuda.current_stream() if torch.cuda.is_available() else None
    for st in list(_side_dirty):
        cur.wait_stream(st)
    _side_dirty.clear()


class _Inline:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _wgrad_scope(direct, device, *tensors):
    return _OnSide(device, *tensors) if (direct and WGRAD_SIDE_STREAM and lib._profile is None) else _Inline()


# Zero-initialised scratch (BN statistic accumulators, reduction buffers): one arena per device, cleared by ONE memset at the
# start of a step (FlatParams.zero_grad) and bump-allocated, instead of ~200 torch.zeros launches per step.  Views are only
# valid within the step that allocated them; outside a training loop (no zero_grad) the arena runs out and torch.zeros serves.
_ZARENA_BYTES = 16 << 20
_zarena = {}


def new_step_scratch(device):
    check_pending_addends()
    st = _zarena.get(device.index)
    if st is None:
        _zarena[device.index] = [torch.zeros(_ZARENA_BYTES, dtype=torch.uint8, device=device), 0]
    else:
        st[0].zero_()
        st[1] = 0


def zeros_scratch(shape, dtype, device):
    st = _zarena.get(device.index)
    n = 1
    for d in (shape if isinstance(shape, (tuple, list)) else (shape,)):
        n *= int(d)
    nbytes = n * torch.empty(0, dtype=dtype).element_size()
    if st is None or st[1] + nbytes > _ZARENA_BYTES:
        return torch.zeros(shape, dtype=dtype, device=device)
    off = st[1]
    st[1] = (off + nbytes + 15) & ~15
    return st[0][off:off + nbytes].view(dtype).view(shape)


def _out_size(i, k, stride, pad, dil):
    return (i + 2 * pad - dil * (k - 1) - 1) // stride + 1


def _tc_ok(x):
    return USE_TENSOR_CORES and x.dtype == torch.bfloat16


def _conv_gather(src, wpacked, bias, n, h, w, c, oh, ow, k, r, s, stride, pad, dil, ups, groups, out_dtype):
    out = empty_act(n, k, oh, ow, out_dtype, src.device)
    g = ConvGeom(n, h, w, c, oh, ow, k, r, s, stride, pad, dil, ups, groups, dtype_code(src),
                 F32 if out_dtype == torch.float32 else BF16)
    lib.note_work(2.0 * n * oh * ow * k * (c // groups) * r * s / (ups * ups), src.element_size() * n * (h * w * c + oh * ow * k),
                  f"simt n{n} {h}x{w}->{oh}x{ow} c{c} k{k} {r}x{s} s{stride} g{groups} u{ups}")
    call("xv2_conv_gather_simt", g, ptr(src), ptr(wpacked), ptr(bias), ptr(out))
    return out


class _Conv2d(torch.autograd.Function):
    """nn.Conv2d (layers.py:92,71; encoder convs unet.py:52) with an optional second source concatenated on channels."""

    @staticmethod
    def forward(ctx, x, x2, weight, bias, stride, pad, dil, groups, want_stats=False):
        out, stats = _Conv2d._forward(ctx, x, x2, weight, bias, stride, pad, dil, groups, want_stats)
        if stats is None:
            stats = torch.empty(0, dtype=torch.float64, device=out.device)
        ctx.mark_non_differentiable(stats)
        ctx.set_materialize_grads(False)  # else autograd launches a zero-fill for the statistics' "gradient" before every backward
        return out, stats

    @staticmethod
    def _forward(ctx, x, x2, weight, bias, stride, pad, dil, groups, want_stats):
        _require_cuda(x)
        x = nhwc(x)
        n, c0, h, w = x.shape
        c1 = 0
        if x2 is not None:
            x2 = nhwc(x2)
            c1 = x2.shape[1]
        k, cg, r, s = weight.shape
        assert (c0 + c1) == cg * groups, "channel mismatch"
        oh, ow = _out_size(h, r, stride, pad, dil), _out_size(w, s, stride, pad, dil)
        ctx.cfg = (stride, pad, dil, groups, c0, c1)
        ctx.save_for_backward(x, x2, weight)
        ctx.has_bias = bias is not None
        ctx.stem = (_tc_ok(x) and stride == 2 and r == 3 and s == 3 and pad == 1 and dil == 1 and c0 == 3 and c1 == 0 and
                    groups == 1 and k in (32, 64) and bias is None)
        if ctx.stem:  # Cin = 3 stride-2 stem conv: direct CUDA-core kernel (K = 27 is no tensor-core shape)
            out = empty_act(n, k, oh, ow, x.dtype, x.device)
            lib.note_work(2.0 * n * oh * ow * k * 27, 2.0 * n * (h * w * 3 + oh * ow * k), f"stem n{n} {h}x{w} k{k}")
            call("xv2_stem_conv_fwd", ptr(x), ptr(_weight_phys(weight)), ptr(out), n, h, w, k)
            return out, None
        use_tc = _tc_ok(x) and stride == 1 and oh == h and ow == w
        if use_tc:
            wp = pack_weight(weight, 0, torch.bfloat16, groups)
            out = empty_act(n, k, h, w, x.dtype, x.device)
            p = TcConv(n, h, w, c0, c1, 0, 0, k, r, s, pad, dil, groups, 0, BF16, 0)
            lib.note_work(2.0 * n * h * w * k * cg * r * s, 2.0 * n * h * w * (c0 + c1 + k) + 2.0 * k * cg * r * s,
                          f"fwd n{n} {h}x{w} c{c0}+{c1} k{k} {r}x{s} g{groups}")
            if want_stats:  # BN statistics fused into the conv epilogue (shapes served by the strip kernel)
                stats = zeros_scratch(2 * k, torch.float64, x.device)
                work = lib._work
                rc = call("xv2_conv_tc", p, ptr(x), ptr(x2), ptr(wp), ptr(bias), ptr(out), ptr(stats), allow_unsupported=True)
                if rc == 0:
                    return out, stats
                lib._work = work
            rc = call("xv2_conv_tc", p, ptr(x), ptr(x2), ptr(wp), ptr(bias), ptr(out), None, allow_unsupported=True)
            if rc == 0:
                return out, None
        src = x if x2 is None else torch.cat((x, x2), 1)  # SIMT path only: plumbing copy
        src = nhwc(src)
        wp = pack_weight(weight, 0, x.dtype, groups)
        return _conv_gather(src, wp, bias, n, h, w, c0 + c1, oh, ow, k, r, s, stride, pad, dil, 1, groups, x.dtype), None

    @staticmethod
    def backward(ctx, dy, _dstats=None):
        if dy is None:
            return (None,) * 9
        x, x2, weight = ctx.saved_tensors
        stride, pad, dil, groups, c0, c1 = ctx.cfg
        dy = nhwc(dy)
        n, _, h, w = x.shape
        k, cg, r, s = weight.shape
        oh, ow = dy.shape[2], dy.shape[3]
        dx = dx2 = dw = db = None
        same = stride == 1 and oh == h and ow == w
        tc = _tc_ok(x) and same
        need_dx = ctx.needs_input_grad[0] or (x2 is not None and ctx.needs_input_grad[1])
        if need_dx:
            pad_t = dil * (r - 1) - pad
            done = False
            if tc:
                wt = pack_weight(weight, 1, torch.bfloat16, groups).view(c0 + c1, -1)
                outs = []
                ok = True
                for lo, cc in ((0, c0), (c0, c1)):
                    if cc == 0:
                        outs.append(None)
                        continue
                    if groups > 1:
                        wsub, kk, gg = wt, k, groups
                    else:
                        wsub, kk, gg = wt[lo:lo + cc], k, 1
                    o = empty_act(n, cc, h, w, x.dtype, x.device)
                    p = TcConv(n, h, w, kk, 0, 0, 0, cc, r, s, pad_t, dil, gg, 0, BF16, 0)
                    lib.note_work(2.0 * n * h * w * cc * (k // gg) * r * s, 2.0 * n * h * w * (k + cc) + 2.0 * cc * (k // gg) * r * s,
                                  f"dgrad n{n} {h}x{w} c{k} k{cc} {r}x{s} g{gg}")
                    rc = call("xv2_conv_tc", p, ptr(dy), None, ptr(wsub), None, ptr(o), None, allow_unsupported=True)
                    if rc != 0:
                        ok = False
                        break
                    outs.append(o)
                if ok:
                    dx, dx2 = outs
                    done = True
            if not done:
                wt = pack_weight(weight, 1, x.dtype, groups)
                full = _conv_gather(dy, wt, None, n, oh, ow, k, h, w, c0 + c1, r, s, 1, pad_t, dil, stride, groups, x.dtype)
                if x2 is None:
                    dx = full
                else:
                    dx, dx2 = nhwc(full[:, :c0]), nhwc(full[:, c0:])
        direct = False
        if ctx.needs_input_grad[2]:
            dw, direct = _grad_buffer(weight)
            done = False
            if tc:
                p = TcConv(n, h, w, c0, c1, 0, 0, k, r, s, pad, dil, groups, 0, BF16, 0)
                lib.note_work(2.0 * n * h * w * k * cg * r * s, 2.0 * n * h * w * (c0 + c1 + k) + 4.0 * k * cg * r * s,
                              f"wgrad n{n} {h}x{w} c{c0}+{c1} k{k} {r}x{s} g{groups}")
                with _wgrad_scope(direct, x.device, x, x2, dy):
                    rc = call("xv2_wgrad_tc", p, ptr(x), ptr(x2), ptr(dy), 0, ptr(dw), allow_unsupported=True)
                done = rc == 0
            if not done and ctx.stem:
                lib.note_work(2.0 * n * oh * ow * k * 27, 2.0 * n * (h * w * 3 + oh * ow * k), f"stem wgrad n{n} {h}x{w} k{k}")
                call("xv2_stem_conv_wgrad", ptr(x), ptr(dy), ptr(dw), n, h, w, k)
                done = True
            if not done:
                src = x if x2 is None else nhwc(torch.cat((x, x2), 1))
                g = ConvGeom(n, h, w, c0 + c1, oh, ow, k, r, s, stride, pad, dil, 1, groups, dtype_code(x), F32)
                call("xv2_conv_wgrad_simt", g, ptr(src), ptr(dy), ptr(dw))
        if ctx.has_bias and ctx.needs_input_grad[3]:
            db = torch.zeros(k, dtype=torch.float32, device=x.device)
            _colsum(dy, db)
        return dx, dx2, (None if direct else dw), db, None, None, None, None, None


def _colsum(t, out):
    n, k, h, w = t.shape
    if k <= 256:
        call("xv2_colsum", ptr(t), n * h * w, k, dtype_code(t), ptr(out))
    else:
        stats = torch.zeros(2 * k, dtype=torch.float64, device=t.device)
        call("xv2_bn_stats", ptr(t), n * h * w, k, dtype_code(t), ptr(stats))
        out.copy_(stats[:k])


def conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, x2=None):
    return _Conv2d.apply(x, x2, weight, bias, stride, padding, dilation, groups, False)[0]


def conv2d_stats(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, x2=None):
    """Convolution that also returns the fp64 [2k] (sum, sum of squares) of its rounded outputs when the kernel could
    fuse them into its epilogue (else None): feeds batch_norm_act(..., stats=...)."""
    out, stats = _Conv2d.apply(x, x2, weight, bias, stride, padding, dilation, groups, True)
    return out, (stats if stats.numel() else None)


class _ConvT2x2(torch.autograd.Function):
    """nn.ConvTranspose2d(k=2, s=2, bias=False) (layers.py:83) as GEMM + pixel-shuffle store."""

    @staticmethod
    def forward(ctx, x, weight):
        _require_cuda(x)
        x = nhwc(x)
        n, cin, h, w = x.shape
        cout = weight.shape[1]
        assert weight.shape[0] == cin and weight.shape[2:] == (2, 2)
        ctx.save_for_backward(x, weight)
        if _tc_ok(x):
            wp = pack_weight(weight, 2, torch.bfloat16)
            out = empty_act(n, cout, 2 * h, 2 * w, x.dtype, x.device)
            p = TcConv(n, h, w, cin, 0, 0, 0, cout, 1, 1, 0, 1, 1, 1, BF16, 0)
            lib.note_work(8.0 * n * h * w * cin * cout, 2.0 * n * h * w * (cin + 4 * cout) + 8.0 * cin * cout,
                          f"convT fwd n{n} {h}x{w} c{cin} k{cout}")
            rc = call("xv2_conv_tc", p, ptr(x), None, ptr(wp), None, ptr(out), None, allow_unsupported=True)
            if rc == 0:
                return out
        wp = pack_weight(weight, 1, x.dtype)  # [cout][1-kh][1-kw][cin]
        return _conv_gather(x, wp, None, n, h, w, cin, 2 * h, 2 * w, cout, 2, 2, 1, 1, 1, 2, 1, x.dtype)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = nhwc(dy)
        n, cin, h, w = x.shape
        cout = weight.shape[1]
        dx = dw = None
        tc = _tc_ok(x)
        if ctx.needs_input_grad[0]:
            done = False
            if tc:
                wp = pack_weight(weight, 0, torch.bfloat16)  # physical [cin][kh][kw][cout] is already the GEMM order
                dx = empty_act(n, cin, h, w, x.dtype, x.device)
                p = TcConv(n, h, w, cout, 0, 0, 0, cin, 2, 2, 0, 1, 1, 2, BF16, 0)
                lib.note_work(8.0 * n * h * w * cin * cout, 2.0 * n * h * w * (cin + 4 * cout) + 8.0 * cin * cout,
                              f"convT dgrad n{n} {h}x{w} c{cout} k{cin}")
                rc = call("xv2_conv_tc", p, ptr(dy), None, ptr(wp), None, ptr(dx), None, allow_unsupported=True)
                done = rc == 0
            if not done:
                wp = pack_weight(weight, 0, x.dtype)
                dx = _conv_gather(dy, wp, None, n, 2 * h, 2 * w, cout, h, w, cin, 2, 2, 2, 0, 1, 1, 1, x.dtype)
        direct = False
        if ctx.needs_input_grad[1]:
            dw, direct = _grad_buffer(weight)
            done = False
            if tc:
                p = TcConv(n, h, w, cin, 0, 0, 0, cout, 2, 2, 0, 1, 1, 1, BF16, 0)
                lib.note_work(8.0 * n * h * w * cin * cout, 2.0 * n * h * w * (cin + 4 * cout) + 16.0 * cin * cout,
                              f"convT wgrad n{n} {h}x{w} c{cin} k{cout}")
                with _wgrad_scope(direct, x.device, x, dy):
                    rc = call("xv2_wgrad_tc", p, ptr(x), None, ptr(dy), 0, ptr(dw), allow_unsupported=True)
                done = rc == 0
            if not done:
                # gradient of the stride-2 conv whose "input" is dy and "output" is x: dw[cin][kh][kw][cout]
                g = ConvGeom(n, 2 * h, 2 * w, cout, h, w, cin, 2, 2, 2, 0, 1, 1, 1, dtype_code(x), F32)
                call("xv2_conv_wgrad_simt", g, ptr(dy), ptr(x), ptr(dw))
        return dx, (None if direct else dw)


def conv_transpose2x2(x, weight):
    return _ConvT2x2.apply(x, weight)


# ---------------------------------------------------------------------------------------------------------------
# batch norm (+ activation, + residual)
# ---------------------------------------------------------------------------------------------------------------
class _BatchNormAct(torch.autograd.Function):
    """nn.BatchNorm2d [+ residual add] [+ ReLU / LeakyReLU(0.01)] in one apply pass (layers.py:93-94, unet.py:52)."""

    @staticmethod
    def forward(ctx, x, residual, gamma, beta, running_mean, running_var, training, momentum, eps, act, stats=None):
        _require_cuda(x)
        x = nhwc(x)
        n, c, h, w = x.shape
        pixels = n * h * w
        dev = x.device
        if residual is not None:
            residual = nhwc(residual)
        coef = torch.empty(4, c, dtype=torch.float32, device=dev)  # mean, invstd, scale, shift
        mean, invstd, scale, shift = coef[0], coef[1], coef[2], coef[3]
        if training:
            if pixels <= 1:
                raise ValueError("Expected more than 1 value per channel when training")  # torch's own BN check
            if stats is None:
                stats = zeros_scratch(2 * c, torch.float64, dev)
                call("xv2_bn_stats", ptr(x), pixels, c, dtype_code(x), ptr(stats))
            y = torch.empty_like(x)
            call("xv2_bn_train_apply", ptr(x), ptr(residual), ptr(y), pixels, c, dtype_code(x), ptr(stats), pixels, ptr(gamma),
                 ptr(beta), ptr(running_mean), ptr(running_var), float(momentum), float(eps), ptr(coef), act)
        else:
            mean.copy_(running_mean)
            invstd.copy_(torch.rsqrt(running_var + eps))
            call("xv2_bn_eval_coeffs", c, ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), float(eps),
                 ptr(scale), ptr(shift))
            y = torch.empty_like(x)
            call("xv2_bn_apply", ptr(x), ptr(residual), ptr(y), pixels, c, dtype_code(x), ptr(scale), ptr(shift), act)
        ctx.save_for_backward(x, residual, gamma, coef)
        ctx.cfg = (training, act)
        ctx.flat_grads = (None, None)
        if all(getattr(p, "_xv2_flat", False) and p.grad is not None and p.grad.dtype == torch.float32 and p.grad.is_contiguous()
               for p in (gamma, beta)):
            ctx.flat_grads = (gamma.grad, beta.grad)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, residual, gamma, coef = ctx.saved_tensors
        training, act = ctx.cfg
        dy2 = _pop_addend(dy)  # second part of the gradient (ops.fork), summed inside the reduce kernel
        dy = nhwc(dy)
        n, c, h, w = x.shape
        pixels = n * h * w
        mean, invstd, scale, shift = coef[0], coef[1], coef[2], coef[3]
        dt = dtype_code(x)
        red = zeros_scratch(2 * c, torch.float64, x.device)
        gg, gb = ctx.flat_grads
        want_dres = residual is not None and ctx.needs_input_grad[1]
        if training and (residual is not None or dy2 is not None) and dt == BF16:
            # residual join: pass 1 writes du = (dy [+ dy2]) * act'(u) once -- it IS the shortcut's gradient -- and pass 2 reads
            # (du, x) only: 7 streaming passes (8 with dy2) instead of 8 plus autograd's 3-pass accumulation of dy + dy2
            du = torch.empty_like(x)
            rc = call("xv2_bn_bwd_reduce_du", ptr(dy), ptr(dy2), ptr(x), ptr(residual), ptr(du), pixels, c, dt, ptr(scale),
                      ptr(shift), ptr(mean), ptr(invstd), act, ptr(red), allow_unsupported=True)
            if rc == 0:
                dx = torch.empty_like(x)
                dgb = None if gg is not None else torch.empty(2, c, dtype=torch.float32, device=x.device)
                call("xv2_bn_bwd_apply", ptr(du), ptr(x), None, ptr(dx), None, pixels, c, dt, ptr(scale), ptr(shift), ptr(mean),
                     ptr(invstd), ptr(gamma), ACT_NONE, ptr(red), pixels, ptr(gg if dgb is None else dgb[0]),
                     ptr(gb if dgb is None else dgb[1]), 1 if dgb is None else 0)
                return (dx, (du if want_dres else None), None if dgb is None else dgb[0], None if dgb is None else dgb[1],
                        None, None, None, None, None, None, None)
        if dy2 is not None:
            dy = nhwc(dy + nhwc(dy2))
        call("xv2_bn_bwd_reduce", ptr(dy), ptr(x), ptr(residual), pixels, c, dt, ptr(scale), ptr(shift), ptr(mean),
             ptr(invstd), act, ptr(red))
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if want_dres else None
        if training and gg is not None:
            # the parameters' own gradient slots in the flat buffer (zeroed once per step): the kernel adds to them, autograd
            # launches no accumulation kernel for gamma / beta
            call("xv2_bn_bwd_apply", ptr(dy), ptr(x), ptr(residual), ptr(dx), ptr(dres), pixels, c, dt, ptr(scale),
                 ptr(shift), ptr(mean), ptr(invstd), ptr(gamma), act, ptr(red), pixels, ptr(gg), ptr(gb), 1)
            return dx, dres, None, None, None, None, None, None, None, None, None
        dgb = torch.empty(2, c, dtype=torch.float32, device=x.device)
        call("xv2_bn_bwd_apply", ptr(dy), ptr(x), ptr(residual), ptr(dx), ptr(dres), pixels, c, dt, ptr(scale),
             ptr(shift), ptr(mean), ptr(invstd), ptr(gamma), act, ptr(red) if training else None, pixels,
             ptr(dgb[0]) if training else None, ptr(dgb[1]) if training else None, 0)
        if not training:
            dgb[1].copy_(red[:c])
            dgb[0].copy_(red[c:])
        return dx, dres, dgb[0], dgb[1], None, None, None, None, None, None, None


# A block output consumed twice by the next residual block (its conv1 and its shortcut) would get its two gradients summed by an
# autograd accumulation kernel (3 passes over the activation).  ops.fork splits the tensor into two aliases; in backward it
# passes the conv1 part on and parks the shortcut part here, keyed by the storage it travels with, for the producer's
# _BatchNormAct.backward to pick up (xv2_bn_bwd_reduce_du sums them while it streams).  check_pending_addends() -- run at every
# step boundary -- fails loudly if a parked part was never collected.
FORK_GRADS = _os.environ.get("XV2_NO_FORK", "0") != "1"
_pending_addends = {}


def _pop_addend(dy):
    hit = _pending_addends.pop(dy.data_ptr(), None) if _pending_addends else None
    return None if hit is None else hit[1]


def check_pending_addends():
    if _pending_addends:
        n = len(_pending_addends)
        _pending_addends.clear()
        raise RuntimeError(f"ops.fork: {n} parked gradient part(s) were never consumed by a BatchNorm backward")


class _Fork(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x), x.view_as(x)

    @staticmethod
    def backward(ctx, da, db):
        if da is None or db is None:
            return da if db is None else db
        da = nhwc(da)
        _pending_addends[da.data_ptr()] = (da, nhwc(db))  # holding `da` keeps its storage from being reused while parked
        return da


def fork(x):
    """Two aliases of a train-mode batch_norm_act output that has NO other consumer (a block output feeding the next block's
    conv1 and shortcut; an encoder stage output feeding the next stage and the decoder); anything else gets (x, x) and
    autograd's own accumulation."""
    if FORK_GRADS and getattr(x, "_xv2_forkable", False) and torch.is_grad_enabled() and x.requires_grad:
        return _Fork.apply(x)
    return x, x


# nn.BatchNorm2d.num_batches_tracked += 1 is one tiny launch per BatchNorm per step (81 at config 2).  Inside `defer_nbt()` (the
# captured training step) the bumps are collected and applied by one multi-tensor add per multiplicity at the end of the forward.
_nbt_pending = None


class defer_nbt:
    def __enter__(self):
        global _nbt_pending
        self.prev, _nbt_pending = _nbt_pending, []
        return self

    def __exit__(self, *exc):
        global _nbt_pending
        pending, _nbt_pending = _nbt_pending, self.prev
        if exc[0] is None and pending:
            count = {}
            for t in pending:
                count.setdefault(id(t), [t, 0])[1] += 1
            for m in sorted({c for _, c in count.values()}):
                torch._foreach_add_([t for t, c in count.values() if c == m], m)
        return False


def _bump_nbt(bn):
    t = bn.num_batches_tracked
    if t is None:
        return
    if _nbt_pending is not None:
        _nbt_pending.append(t)
    else:
        t += 1


def _flat_grads(*params):
    """The parameters' own gradient slots in the flat buffer (optim.FlatParams; zeroed once per step) when ALL of them have one:
    kernels add into them directly and autograd launches no accumulation kernel."""
    if all(p is not None and getattr(p, "_xv2_flat", False) and p.grad is not None and p.grad.dtype == torch.float32 and
           p.grad.is_contiguous() for p in params):
        return tuple(p.grad for p in params)
    return None


def batch_norm_act(x, bn, act=ACT_NONE, residual=None, stats=None):
    """Applies the nn.BatchNorm2d module `bn` (its parameters / buffers / training flag) followed by `act`."""
    if bn.training and bn.track_running_stats:
        _bump_nbt(bn)
    use_batch = bn.training or not bn.track_running_stats
    y = _BatchNormAct.apply(x, residual, bn.weight, bn.bias, bn.running_mean, bn.running_var, use_batch,
                            bn.momentum if bn.momentum is not None else 0.1, bn.eps, act, stats if use_batch else None)
    if use_batch and y.dtype == torch.bfloat16:
        y._xv2_forkable = True  # see ops.fork
    return y


def _eval_coeffs(bn):
    """(scale, shift) of an eval-mode BatchNorm, cached on the module until its parameters / running statistics change."""
    stamp = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version, _pack_epoch,
             bn.weight.data_ptr())
    hit = bn.__dict__.get("_xv2_eval")
    if hit is not None and hit[0] == stamp:
        return hit[1]
    c = bn.num_features
    coef = torch.empty(2, c, dtype=torch.float32, device=bn.weight.device)
    call("xv2_bn_eval_coeffs", c, ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean), ptr(bn.running_var), float(bn.eps),
         ptr(coef[0]), ptr(coef[1]))
    bn.__dict__["_xv2_eval"] = (stamp, coef)
    return coef


def _conv_bnact_inference(x, conv, bn, act, x2):
    """Inference (no autograd, eval-mode BN): BatchNorm + activation folded into the conv epilogue -- no BN pass at all.
    Returns None when the tensor-core kernels do not serve the shape."""
    stride, pad, dil, groups = conv.stride[0], conv.padding[0], conv.dilation[0], conv.groups
    if not (_tc_ok(x) and stride == 1 and conv.bias is None):
        return None
    x = nhwc(x)
    n, c0, h, w = x.shape
    c1 = 0
    if x2 is not None:
        x2 = nhwc(x2)
        c1 = x2.shape[1]
    k, cg, r, s = conv.weight.shape
    if _out_size(h, r, 1, pad, dil) != h or _out_size(w, s, 1, pad, dil) != w:
        return None
    coef = _eval_coeffs(bn)
    wp = pack_weight(conv.weight, 0, torch.bfloat16, groups)
    out = empty_act(n, k, h, w, x.dtype, x.device)
    p = TcConv(n, h, w, c0, c1, 0, 0, k, r, s, pad, dil, groups, 0, BF16, 0)
    lib.note_work(2.0 * n * h * w * k * cg * r * s, 2.0 * n * h * w * (c0 + c1 + k) + 2.0 * k * cg * r * s,
                  f"fwd+bn n{n} {h}x{w} c{c0}+{c1} k{k} {r}x{s} g{groups}")
    rc = call("xv2_conv_tc_bnact", p, ptr(x), ptr(x2), ptr(wp), ptr(coef[0]), ptr(coef[1]), act, ptr(out), allow_unsupported=True)
    return out if rc == 0 else None


def conv_bn_act(x, conv, bn, act=ACT_NONE, x2=None, residual=None):
    """conv -> BatchNorm -> activation (+ residual) for an nn.Conv2d / nn.BatchNorm2d parameter pair; in training the BN
    statistics come out of the conv epilogue when the kernel supports it; at inference (eval-mode BN, autograd off) the
    BatchNorm and the activation are folded into the conv epilogue."""
    args = (x, conv.weight, conv.bias, conv.stride[0], conv.padding[0], conv.dilation[0], conv.groups, x2)
    if not bn.training and bn.track_running_stats and residual is None and not torch.is_grad_enabled():
        out = _conv_bnact_inference(x, conv, bn, act, x2)
        if out is not None:
            return out
    if bn.training or not bn.track_running_stats:
        out, stats = conv2d_stats(*args)
    else:
        out, stats = conv2d(*args), None
    return batch_norm_act(out, bn, act, residual, stats)


def _flat_grad_pair(gamma, beta):
    if all(getattr(p, "_xv2_flat", False) and p.grad is not None and p.grad.dtype == torch.float32 and p.grad.is_contiguous()
           for p in (gamma, beta)):
        return gamma.grad, beta.grad
    return None, None


class _BNActHead(torch.autograd.Function):
    """Train-mode BatchNorm + activation of the LAST decoder ConvLayer fused with the 1x1 output head (layers.py:96-100 then
    :180-183): the full-resolution activation between them is never written (xview2_b200/csrc/fused_tail.cu)."""

    @staticmethod
    def forward(ctx, z, stats, gamma, beta, running_mean, running_var, momentum, eps, act, head_w, head_b):
        z = nhwc(z)
        n, c, h, w = z.shape
        pixels = n * h * w
        dev = z.device
        ncls = head_w.shape[0]
        coef = torch.empty(4, c, dtype=torch.float32, device=dev)  # mean, invstd, scale, shift
        call("xv2_bn_finalize", ptr(stats), pixels, c, ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var),
             float(momentum), float(eps), ptr(coef[0]), ptr(coef[1]), ptr(coef[2]), ptr(coef[3]))
        w2 = head_w.reshape(ncls, c).contiguous()
        logits = empty_act(n, ncls, h, w, torch.float32, dev)
        lib.note_work(0.0, 2.0 * pixels * c + 4.0 * pixels * ncls, f"bn+act+head fwd n{n} {h}x{w} c{c} ncls{ncls}")
        call("xv2_bnact_head_fwd", ptr(z), pixels, c, ptr(coef[2]), ptr(coef[3]), act, ptr(w2), ptr(head_b), ncls, ptr(logits))
        ctx.save_for_backward(z, gamma, coef, w2)
        ctx.cfg = (act, head_w.shape, head_b is not None)
        ctx.flat_grads = _flat_grad_pair(gamma, beta)
        ctx.head_flat = _flat_grads(head_w, head_b) if head_b is not None else None
        return logits

    @staticmethod
    def backward(ctx, dl):
        z, gamma, coef, w2 = ctx.saved_tensors
        act, wshape, has_bias = ctx.cfg
        n, c, h, w = z.shape
        pixels = n * h * w
        ncls = w2.shape[0]
        dev = z.device
        dl = nhwc(dl.float())
        red = zeros_scratch(2 * c, torch.float64, dev)
        hflat = ctx.head_flat  # head weight / bias slots of the flat gradient buffer: the reduce kernel atomically adds into them
        dhw = hflat[0] if hflat is not None else zeros_scratch((ncls, c), torch.float32, dev)
        dhb = hflat[1] if hflat is not None else zeros_scratch(ncls, torch.float32, dev)
        lib.note_work(0.0, 2.0 * pixels * c + 4.0 * pixels * ncls, f"bn+act+head bwd reduce n{n} {h}x{w} c{c}")
        call("xv2_bnact_head_bwd_reduce", ptr(z), ptr(dl), pixels, c, ptr(coef[2]), ptr(coef[3]), ptr(coef[0]), ptr(coef[1]), act,
             ptr(w2), ncls, ptr(red), ptr(dhw), ptr(dhb))
        dz = torch.empty_like(z)
        gg, gb = ctx.flat_grads
        dgb = None if gg is not None else torch.empty(2, c, dtype=torch.float32, device=dev)
        lib.note_work(0.0, 4.0 * pixels * c + 4.0 * pixels * ncls, f"bn+act+head bwd apply n{n} {h}x{w} c{c}")
        call("xv2_bnact_head_bwd_apply", ptr(z), ptr(dl), ptr(dz), pixels, c, ptr(coef[2]), ptr(coef[3]), ptr(coef[0]), ptr(coef[1]),
             ptr(gamma), act, ptr(w2), ncls, ptr(red), pixels, ptr(gg if gg is not None else dgb[0]),
             ptr(gb if gb is not None else dgb[1]), 1 if gg is not None else 0)
        if hflat is not None:
            return (dz, None, None if gg is not None else dgb[0], None if gg is not None else dgb[1], None, None, None, None, None,
                    None, None)
        return (dz, None, None if gg is not None else dgb[0], None if gg is not None else dgb[1], None, None, None, None, None,
                dhw.clone().reshape(wshape), dhb.clone() if has_bias else None)


class DeferredBNAct:
    """Raw conv output `z` whose train-mode BatchNorm + activation has not been applied yet: the consumer either fuses it
    (ops.bnact_head) or calls materialise()."""

    def __init__(self, z, stats, bn, act):
        self.z, self.stats, self.bn, self.act = z, stats, bn, act

    def materialise(self):
        return batch_norm_act(self.z, self.bn, self.act, None, self.stats)


def conv_bn_act_deferred(x, conv, bn, act, x2=None):
    """conv now, BatchNorm + activation later (only while the BatchNorm uses batch statistics on the bf16 path)."""
    if not (FUSE_TAIL and bn.training and bn.track_running_stats and _tc_ok(x) and conv.out_channels in (32, 64)):
        return conv_bn_act(x, conv, bn, act, x2=x2)
    out, stats = conv2d_stats(x, conv.weight, conv.bias, conv.stride[0], conv.padding[0], conv.dilation[0], conv.groups, x2)
    return DeferredBNAct(out, stats, bn, act)


def bnact_head(deferred, head_w, head_b):
    """Consumes a DeferredBNAct with the fused BatchNorm + activation + 1x1 head kernels."""
    bn, z = deferred.bn, nhwc(deferred.z)
    n, c, h, w = z.shape
    stats = deferred.stats
    if stats is None:
        stats = zeros_scratch(2 * c, torch.float64, z.device)
        call("xv2_bn_stats", ptr(z), n * h * w, c, dtype_code(z), ptr(stats))
    _bump_nbt(bn)
    return _BNActHead.apply(z, stats, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                            bn.momentum if bn.momentum is not None else 0.1, bn.eps, deferred.act, head_w, head_b)


# ---------------------------------------------------------------------------------------------------------------
# pooling
# ---------------------------------------------------------------------------------------------------------------
def _pool_out(i, k, s, p, ceil_mode):
    if ceil_mode:
        o = -(-(i + 2 * p - k) // s) + 1
        if (o - 1) * s >= i + p:
            o -= 1
        return o
    return (i + 2 * p - k) // s + 1


class _MaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, stride, pad):
        _require_cuda(x)
        x = nhwc(x)
        n, c, h, w = x.shape
        oh, ow = _pool_out(h, k, stride, pad, False), _pool_out(w, k, stride, pad, False)
        y = empty_act(n, c, oh, ow, x.dtype, x.device)
        idx = torch.empty((n, oh, ow, c), dtype=torch.uint8, device=x.device) if ctx.needs_input_grad[0] else None
        call("xv2_maxpool_fwd", ptr(x), ptr(y), ptr(idx), n, h, w, c, oh, ow, k, stride, pad, dtype_code(x))
        ctx.save_for_backward(idx)
        ctx.cfg = (k, stride, pad, oh, ow, n, c, h, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        k, stride, pad, oh, ow, n, c, h, w = ctx.cfg
        dy = nhwc(dy)
        dx = empty_act(n, c, h, w, dy.dtype, dy.device)
        call("xv2_maxpool_bwd", ptr(idx), ptr(dy), ptr(dx), n, h, w, c, oh, ow, k, stride, pad, dtype_code(dy))
        return dx, None, None, None


class _AvgPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, stride, pad, ceil_mode, count_include_pad):
        _require_cuda(x)
        x = nhwc(x)
        n, c, h, w = x.shape
        oh, ow = _pool_out(h, k, stride, pad, ceil_mode), _pool_out(w, k, stride, pad, ceil_mode)
        y = empty_act(n, c, oh, ow, x.dtype, x.device)
        call("xv2_avgpool_fwd", ptr(x), ptr(y), n, h, w, c, oh, ow, k, stride, pad, int(count_include_pad), dtype_code(x))
        ctx.cfg = (k, stride, pad, int(count_include_pad), n, c, h, w, oh, ow)
        return y

    @staticmethod
    def backward(ctx, dy):
        k, stride, pad, cip, n, c, h, w, oh, ow = ctx.cfg
        dy = nhwc(dy)
        dx = empty_act(n, c, h, w, dy.dtype, dy.device)
        call("xv2_avgpool_bwd", ptr(dy), ptr(dx), n, h, w, c, oh, ow, k, stride, pad, cip, dtype_code(dy))
        return dx, None, None, None, None, None


def max_pool2d(x, k, stride, pad):
    return _MaxPool.apply(x, k, stride, pad)


def avg_pool2d(x, k, stride, pad=0, ceil_mode=False, count_include_pad=True):
    if k == 1 and stride == 1:
        return x
    return _AvgPool.apply(x, k, stride, pad, ceil_mode, count_include_pad)


class _AdaptiveAvgPool(torch.autograd.Function):
    """nn.AdaptiveAvgPool2d(bins) (PPM, layers.py:12-21)."""

    @staticmethod
    def forward(ctx, x, bins):
        _require_cuda(x)
        x = nhwc(x)
        n, c, h, w = x.shape
        y = empty_act(n, c, bins, bins, x.dtype, x.device)
        call("xv2_adaptive_avgpool_fwd", ptr(x), ptr(y), n, h, w, c, bins, dtype_code(x))
        ctx.cfg = (n, c, h, w, bins)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w, bins = ctx.cfg
        dy = nhwc(dy)
        dx = empty_act(n, c, h, w, dy.dtype, dy.device)
        call("xv2_adaptive_avgpool_bwd", ptr(dy), ptr(dx), n, h, w, c, bins, dtype_code(dy))
        return dx, None


def adaptive_avg_pool2d(x, bins):
    return _AdaptiveAvgPool.apply(x, bins)


class _Bilinear(torch.autograd.Function):
    """F.interpolate(mode="bilinear", align_corners=True) (layers.py:27,154,188)."""

    @staticmethod
    def forward(ctx, x, oh, ow):
        _require_cuda(x)
        x = nhwc(x)
        n, c, h, w = x.shape
        y = empty_act(n, c, oh, ow, x.dtype, x.device)
        call("xv2_bilinear_fwd", ptr(x), ptr(y), n, h, w, c, oh, ow, dtype_code(x))
        ctx.cfg = (n, c, h, w, oh, ow)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w, oh, ow = ctx.cfg
        dy = nhwc(dy)
        acc = torch.zeros((n, c, h, w), dtype=torch.float32, device=dy.device).contiguous(memory_format=CL)
        call("xv2_bilinear_bwd", ptr(dy), ptr(acc), n, h, w, c, oh, ow, dtype_code(dy))
        return (acc if dy.dtype == torch.float32 else cast(acc, dy.dtype)), None, None


def bilinear(x, size):
    oh, ow = (size, size) if isinstance(size, int) else size
    return _Bilinear.apply(x, int(oh), int(ow))


# ---------------------------------------------------------------------------------------------------------------
# split attention (ResNeSt SplAtConv2d tail: radix-sum -> GAP -> fc1 -> bn1 -> relu -> fc2 -> r-softmax -> combine)
# ---------------------------------------------------------------------------------------------------------------
def _fc(x2d, w2d, bias):
    """[n][c] x [k][c]^T + bias -> [n][k], fp32, through the SIMT gather conv on a 1x1 'image'."""
    n, c = x2d.shape
    k = w2d.shape[0]
    out = torch.empty((n, k), dtype=torch.float32, device=x2d.device)
    if c % 4 == 0:
        call("xv2_fc_fwd", ptr(x2d), ptr(w2d), ptr(bias), ptr(out), n, c, k)
        return out
    g = ConvGeom(n, 1, 1, c, 1, 1, k, 1, 1, 1, 0, 1, 1, 1, F32, F32)
    call("xv2_conv_gather_simt", g, ptr(x2d), ptr(w2d), ptr(bias), ptr(out))
    return out


def _fc_wgrad(x2d, dy2d):
    n, c = x2d.shape
    k = dy2d.shape[1]
    if c % 4 == 0:
        dw = torch.empty((k, c), dtype=torch.float32, device=x2d.device)
        db = torch.empty(k, dtype=torch.float32, device=x2d.device)
        call("xv2_fc_wgrad", ptr(x2d), ptr(dy2d), ptr(dw), ptr(db), n, c, k)
        return dw, db
    dw = torch.zeros((k, c), dtype=torch.float32, device=x2d.device)
    g = ConvGeom(n, 1, 1, c, 1, 1, k, 1, 1, 1, 0, 1, 1, 1, F32, F32)
    call("xv2_conv_wgrad_simt", g, ptr(x2d), ptr(dy2d), ptr(dw))
    db = torch.zeros(k, dtype=torch.float32, device=x2d.device)
    if k <= 256:
        call("xv2_colsum", ptr(dy2d), n, k, F32, ptr(db))
    else:
        st = torch.zeros(2 * k, dtype=torch.float64, device=x2d.device)
        call("xv2_bn_stats", ptr(dy2d), n, k, F32, ptr(st))
        db.copy_(st[:k])
    return dw, db


class _SplitAttention(torch.autograd.Function):
    """Whole split-attention tail as one tape node so that d(x) is written in a single pass.

    forward : gap = mean_hw(x0 + x1); a = r-softmax(fc2(relu(bn1(fc1(gap))))); out = a0*x0 + a1*x1
    backward: datt = <dout, x_r>_hw  ->  FC/BN chain  ->  dgap;  dx_r = a_r*dout + dgap/hw
    The FC/BN chain on the [n][C] vectors is fused into 2 forward + 3 backward launches (xv2_splat_fc_fwd / _bwd).
    """

    @staticmethod
    def forward(ctx, x, w1, b1, gamma, beta, rmean, rvar, w2, b2, training, momentum, eps):
        _require_cuda(x)
        x = nhwc(x)
        n, c2, h, w = x.shape
        c = c2 // 2
        inter = w1.shape[0]
        dev = x.device
        if training and n <= 1:
            raise ValueError("Expected more than 1 value per channel when training")
        if n > 32:
            raise lib.Xv2Error("split attention serves batches of at most 32 tiles per GPU")
        gap = torch.empty((n, c), dtype=torch.float32, device=dev)
        call("xv2_splat_gap", ptr(x), ptr(gap), n, h * w, c, dtype_code(x))
        w1m, w2m = w1.reshape(inter, c).contiguous(), w2.reshape(c2, inter).contiguous()
        z1 = torch.empty((n, inter), dtype=torch.float32, device=dev)
        a1 = torch.empty_like(z1)
        coef = torch.empty(4, inter, dtype=torch.float32, device=dev)
        att = torch.empty((n, c2), dtype=torch.float32, device=dev)
        call("xv2_splat_fc_fwd", ptr(gap), ptr(w1m), ptr(b1), ptr(gamma), ptr(beta), ptr(rmean), ptr(rvar), float(momentum),
             float(eps), int(bool(training)), ptr(w2m), ptr(b2), ptr(z1), ptr(a1), ptr(coef), ptr(att), n, c, inter)
        out = empty_act(n, c, h, w, x.dtype, dev)
        call("xv2_splat_combine", ptr(x), ptr(att), ptr(out), n, h * w, c, dtype_code(x))
        ctx.save_for_backward(x, att, gap, z1, a1, coef, w1, w2, gamma)
        ctx.training = training
        return out

    @staticmethod
    def backward(ctx, dout):
        x, att, gap, z1, a1, coef, w1, w2, gamma = ctx.saved_tensors
        training = ctx.training
        dout = nhwc(dout)
        n, c2, h, w = x.shape
        c = c2 // 2
        inter = w1.shape[0]
        dev = x.device
        datt = zeros_scratch((n, c2), torch.float32, dev)
        call("xv2_splat_bwd_att", ptr(x), ptr(dout), ptr(datt), n, h * w, c, dtype_code(x))
        w2t = pack_weight(w2, 1, torch.float32)  # [inter][2c]
        w1t = pack_weight(w1, 1, torch.float32)  # [c][inter]
        scratch = torch.empty(n * (c2 + inter), dtype=torch.float32, device=dev)
        dw2 = torch.empty((c2, inter), dtype=torch.float32, device=dev)
        dw1 = torch.empty((inter, c), dtype=torch.float32, device=dev)
        small = torch.empty(c2 + 3 * inter, dtype=torch.float32, device=dev)
        db2, db1, dgamma, dbeta = small[:c2], small[c2:c2 + inter], small[c2 + inter:c2 + 2 * inter], small[c2 + 2 * inter:]
        dgap = torch.empty((n, c), dtype=torch.float32, device=dev)
        call("xv2_splat_fc_bwd", ptr(att), ptr(datt), ptr(a1), ptr(z1), ptr(coef), ptr(gamma), ptr(gap), ptr(w2t), ptr(w1t),
             int(bool(training)), ptr(scratch), ptr(scratch[n * c2:]), ptr(dw2), ptr(db2), ptr(dw1), ptr(db1), ptr(dgamma),
             ptr(dbeta), ptr(dgap), n, c, inter)
        dx = torch.empty_like(x)
        call("xv2_splat_bwd_x", ptr(dout), ptr(att), ptr(dgap), ptr(dx), n, h * w, c, dtype_code(x))
        return (dx, dw1.reshape(w1.shape), db1, dgamma, dbeta, None, None, dw2.reshape(w2.shape), db2, None, None, None)


class _BnSplitAttention(torch.autograd.Function):
    """bn0 (batch statistics) + ReLU + the whole split-attention tail on the RAW radix-conv output z: the post-BN activation
    is never materialised (xview2_b200/csrc/splat_fused.cu).  Same result as batch_norm_act(z, bn0, RELU) -> _SplitAttention."""

    @staticmethod
    def forward(ctx, z, stats, g0, b0, rm0, rv0, mom0, eps0, w1, b1, gamma, beta, rmean, rvar, w2, b2, training, momentum, eps):
        z = nhwc(z)
        n, c2, h, w = z.shape
        c = c2 // 2
        hw = h * w
        inter = w1.shape[0]
        dev = z.device
        coef0 = torch.empty(4, c2, dtype=torch.float32, device=dev)  # mean, invstd, scale, shift of bn0
        gap = torch.empty((n, c), dtype=torch.float32, device=dev)
        gap_acc = zeros_scratch((n, c), torch.float64, dev)  # fp64 accumulator, zero-filled with the step's arena (no memset node)
        lib.note_work(0.0, 2.0 * n * hw * c2, f"splat bn+gap n{n} {h}x{w} c{c}")
        # bn0's finalize step (coefficients, running statistics) is folded into the GAP kernel's prologue
        call("xv2_splat_bn_gap_fin", ptr(z), ptr(stats), n * hw, ptr(g0), ptr(b0), ptr(rm0), ptr(rv0), float(mom0), float(eps0),
             ptr(coef0), ptr(gap), ptr(gap_acc), n, hw, c)
        w1m, w2m = w1.reshape(inter, c).contiguous(), w2.reshape(c2, inter).contiguous()
        z1 = torch.empty((n, inter), dtype=torch.float32, device=dev)
        a1 = torch.empty_like(z1)
        coef = torch.empty(4, inter, dtype=torch.float32, device=dev)
        att = torch.empty((n, c2), dtype=torch.float32, device=dev)
        call("xv2_splat_fc_fwd", ptr(gap), ptr(w1m), ptr(b1), ptr(gamma), ptr(beta), ptr(rmean), ptr(rvar), float(momentum),
             float(eps), int(bool(training)), ptr(w2m), ptr(b2), ptr(z1), ptr(a1), ptr(coef), ptr(att), n, c, inter)
        out = empty_act(n, c, h, w, z.dtype, dev)
        lib.note_work(0.0, 2.0 * n * hw * (c2 + c), f"splat bn+combine n{n} {h}x{w} c{c}")
        call("xv2_splat_bn_combine", ptr(z), ptr(coef0[2]), ptr(coef0[3]), ptr(att), ptr(out), n, hw, c)
        ctx.save_for_backward(z, att, gap, z1, a1, coef, coef0, w1, w2, gamma, g0)
        ctx.training = training
        ctx.flat_grads = _flat_grad_pair(g0, b0)
        ctx.fc_flat = _flat_grads(w1, b1, gamma, beta, w2, b2)
        return out

    @staticmethod
    def backward(ctx, dout):
        z, att, gap, z1, a1, coef, coef0, w1, w2, gamma, g0 = ctx.saved_tensors
        training = ctx.training
        dout = nhwc(dout)
        n, c2, h, w = z.shape
        c = c2 // 2
        hw = h * w
        inter = w1.shape[0]
        dev = z.device
        part = zeros_scratch((4, n, c2), torch.float64, dev)
        lib.note_work(0.0, 2.0 * n * hw * (c2 + c), f"splat bn bwd partials n{n} {h}x{w} c{c}")
        call("xv2_splat_bn_bwd_partials", ptr(z), ptr(dout), ptr(coef0[2]), ptr(coef0[3]), ptr(part), n, hw, c)
        w2t = pack_weight(w2, 1, torch.float32)  # [inter][2c]
        w1t = pack_weight(w1, 1, torch.float32)  # [c][inter]
        scratch = torch.empty(n * (c2 + inter), dtype=torch.float32, device=dev)
        fcf = ctx.fc_flat
        if fcf is not None:  # the FC parameters' own slots in the flat gradient buffer: the kernels add to them
            dw1, db1, dgamma, dbeta, dw2, db2 = fcf
        else:
            dw2 = torch.empty((c2, inter), dtype=torch.float32, device=dev)
            dw1 = torch.empty((inter, c), dtype=torch.float32, device=dev)
            small = torch.empty(c2 + 3 * inter, dtype=torch.float32, device=dev)
            db2, db1, dgamma, dbeta = small[:c2], small[c2:c2 + inter], small[c2 + inter:c2 + 2 * inter], small[c2 + 2 * inter:]
        dgap = torch.empty((n, c), dtype=torch.float32, device=dev)
        red = torch.empty(2 * c2, dtype=torch.float64, device=dev)
        # datt (from the partial sums) and the two bn0 reductions are computed inside the FC-chain kernels
        call("xv2_splat_fc_bwd_fused", ptr(att), ptr(part), ptr(coef0[2]), ptr(coef0[3]), ptr(coef0[0]), ptr(coef0[1]), hw,
             ptr(a1), ptr(z1), ptr(coef), ptr(gamma), ptr(gap), ptr(w2t), ptr(w1t), int(bool(training)), ptr(scratch),
             ptr(scratch[n * c2:]), ptr(dw2), ptr(db2), ptr(dw1), ptr(db1), ptr(dgamma), ptr(dbeta), ptr(dgap), ptr(red),
             1 if fcf is not None else 0, n, c, inter)
        dz = torch.empty_like(z)
        gg, gb = ctx.flat_grads
        dgb0 = None if gg is not None else torch.empty(2, c2, dtype=torch.float32, device=dev)
        lib.note_work(0.0, 2.0 * n * hw * (2 * c2 + c), f"splat bn bwd apply n{n} {h}x{w} c{c}")
        call("xv2_splat_bn_bwd_apply", ptr(z), ptr(dout), ptr(att), ptr(dgap), ptr(coef0[2]), ptr(coef0[3]), ptr(coef0[0]),
             ptr(coef0[1]), ptr(g0), ptr(red), ptr(dz), ptr(gg if gg is not None else dgb0[0]),
             ptr(gb if gb is not None else dgb0[1]), 1 if gg is not None else 0, n, hw, c)
        g0b0 = (None, None) if gg is not None else (dgb0[0], dgb0[1])
        if fcf is not None:
            return (dz, None, *g0b0, None, None, None, None, None, None, None, None, None, None, None, None, None, None, None)
        return (dz, None, *g0b0, None, None, None, None,
                dw1.reshape(w1.shape), db1, dgamma, dbeta, None, None, dw2.reshape(w2.shape), db2, None, None, None)


def conv_bn_split_attention(x, conv, bn0, fc1, bn1, fc2):
    """SplAtConv2d body: radix conv -> bn0 -> ReLU -> split attention.  With batch statistics on the bf16 path the BatchNorm +
    ReLU are folded into the split-attention kernels (the 2C-channel activation is never written); otherwise the unfused chain."""
    n = x.shape[0]
    c2 = conv.out_channels
    vec = c2 // 16  # channel vectors of one radix half
    fusable = (FUSE_SPLAT_BN and bn0.training and bn0.track_running_stats and bn1.training and _tc_ok(x) and c2 % 16 == 0 and vec >= 1 and
               vec <= 128 and (vec & (vec - 1)) == 0 and 1 < n <= 32)
    if not fusable:
        return split_attention(conv_bn_act(x, conv, bn0, ACT_RELU), fc1, bn1, fc2)
    z, stats = conv2d_stats(x, conv.weight, conv.bias, conv.stride[0], conv.padding[0], conv.dilation[0], conv.groups)
    z = nhwc(z)
    if stats is None:
        stats = zeros_scratch(2 * c2, torch.float64, z.device)
        call("xv2_bn_stats", ptr(z), z.shape[0] * z.shape[2] * z.shape[3], c2, dtype_code(z), ptr(stats))
    for bn in (bn0, bn1):
        _bump_nbt(bn)
    return _BnSplitAttention.apply(z, stats, bn0.weight, bn0.bias, bn0.running_mean, bn0.running_var,
                                   bn0.momentum if bn0.momentum is not None else 0.1, bn0.eps, fc1.weight, fc1.bias,
                                   bn1.weight, bn1.bias, bn1.running_mean, bn1.running_var, fc2.weight, fc2.bias, True,
                                   bn1.momentum, bn1.eps)


def split_attention(x, fc1, bn1, fc2):
    """x: (n, 2c, h, w) after bn0+relu; fc1/fc2: nn.Conv2d 1x1 with bias; bn1: nn.BatchNorm2d on (n, inter, 1, 1)."""
    if bn1.training:
        _bump_nbt(bn1)
    return _SplitAttention.apply(x, fc1.weight, fc1.bias, bn1.weight, bn1.bias, bn1.running_mean, bn1.running_var,
                                 fc2.weight, fc2.bias, bn1.training, bn1.momentum, bn1.eps)


# ---------------------------------------------------------------------------------------------------------------
# element-wise
# ---------------------------------------------------------------------------------------------------------------
class _AddAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, act):
        a, b = nhwc(a), nhwc(b)
        y = torch.empty_like(a)
        call("xv2_add_act", ptr(a), ptr(b), ptr(y), a.numel(), dtype_code(a), act)
        ctx.act = act
        if act != ACT_NONE:
            ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        if ctx.act == ACT_NONE:
            return dy, dy, None
        (y,) = ctx.saved_tensors
        dy = nhwc(dy)
        dx = torch.empty_like(y)
        call("xv2_act_bwd", ptr(dy), ptr(y), ptr(dx), y.numel(), dtype_code(y), ctx.act)
        return dx, dx, None


def add_act(a, b, act=ACT_NONE):
    return _AddAct.apply(a, b, act)


class _Gate(torch.autograd.Function):
    """skip * sigmoid(psi)  (layers.py:165-166); psi has one channel."""

    @staticmethod
    def forward(ctx, skip, psi):
        skip, psi = nhwc(skip), nhwc(psi)
        n, c, h, w = skip.shape
        out = torch.empty_like(skip)
        call("xv2_gate_fwd", ptr(skip), ptr(psi), ptr(out), n * h * w, c, dtype_code(skip))
        ctx.save_for_backward(skip, psi)
        return out

    @staticmethod
    def backward(ctx, dout):
        skip, psi = ctx.saved_tensors
        dout = nhwc(dout)
        n, c, h, w = skip.shape
        dskip, dpsi = torch.empty_like(skip), torch.empty_like(psi)
        call("xv2_gate_bwd", ptr(dout), ptr(skip), ptr(psi), ptr(dskip), ptr(dpsi), n * h * w, c, dtype_code(skip))
        return dskip, dpsi


def gate(skip, psi):
    return _Gate.apply(skip, psi)


class _Flip(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flip_h, flip_w):
        x = nhwc(x)
        n, c, h, w = x.shape
        y = torch.empty_like(x)
        call("xv2_flip", ptr(x), ptr(y), n, h, w, c, int(flip_h), int(flip_w), dtype_code(x))
        ctx.cfg = (flip_h, flip_w)
        return y

    @staticmethod
    def backward(ctx, dy):
        return _Flip.apply(dy, *ctx.cfg), None, None


def flip(x, dims):
    return _Flip.apply(x, 2 in dims, 3 in dims)


class _Cast(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        _require_cuda(x)
        x = nhwc(x)
        ctx.src_dtype = x.dtype
        if x.dtype == dtype:
            return x
        y = torch.empty_like(x, dtype=dtype)
        call("xv2_cast", ptr(x), dtype_code(x), ptr(y), dtype_code(y), x.numel())
        return y

    @staticmethod
    def backward(ctx, dy):
        return _Cast.apply(dy, ctx.src_dtype), None


def cast(x, dtype):
    return _Cast.apply(x, dtype)


# ---------------------------------------------------------------------------------------------------------------
# output head, loss, metric, post-process, loader
# ---------------------------------------------------------------------------------------------------------------
class _Head(torch.autograd.Function):
    """1x1 conv + bias to n_class fp32 logits (layers.py:180)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        _require_cuda(x)
        x = nhwc(x)
        n, c, h, w = x.shape
        ncls = weight.shape[0]
        w2 = weight.reshape(ncls, c).contiguous()
        logits = empty_act(n, ncls, h, w, torch.float32, x.device)
        call("xv2_head_fwd", ptr(x), ptr(w2), ptr(bias), ptr(logits), n * h * w, c, ncls, dtype_code(x))
        ctx.save_for_backward(x, w2)
        ctx.wshape = weight.shape
        return logits

    @staticmethod
    def backward(ctx, dl):
        x, w2 = ctx.saved_tensors
        n, c, h, w = x.shape
        ncls = w2.shape[0]
        dl = nhwc(dl.float())
        dx = torch.empty_like(x)
        dw = torch.zeros_like(w2)
        db = torch.zeros(ncls, dtype=torch.float32, device=x.device)
        call("xv2_head_bwd", ptr(x), ptr(w2), ptr(dl), ptr(dx), ptr(dw), ptr(db), n * h * w, c, ncls, dtype_code(x))
        return dx, dw.reshape(ctx.wshape), db


def head(x, weight, bias):
    return _Head.apply(x, weight, bias)


class _SegLoss(torch.autograd.Function):
    """Sum of dice / focal / ce terms (loss.py:98-101) scaled by `weight`, labels sampled with stride `lstride`."""

    @staticmethod
    def forward(ctx, logits, labels, terms, post, weight, lstride):
        _require_cuda(logits)
        logits = nhwc(logits.float())
        n, ncls, h, w = logits.shape
        labels = labels.contiguous()
        assert labels.dtype == torch.uint8 and labels.shape == (n, h * lstride, w * lstride)
        dev = logits.device
        sums = torch.zeros(3 * ncls + 3, dtype=torch.float64, device=dev)
        call("xv2_loss_partials", ptr(logits), ptr(labels), n, h, w, ncls, lstride, int(post), ptr(sums))
        loss = torch.zeros(1, dtype=torch.float32, device=dev)
        coef = torch.empty(2 * ncls + 4, dtype=torch.float32, device=dev)
        call("xv2_loss_finalize", ptr(sums), ncls, terms, float(weight), ptr(loss), ptr(coef))
        ctx.save_for_backward(logits, labels, coef)
        ctx.cfg = (terms, int(post), lstride)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        logits, labels, coef = ctx.saved_tensors
        terms, post, lstride = ctx.cfg
        n, ncls, h, w = logits.shape
        dl = torch.empty_like(logits)
        up = dloss.reshape(1).float().contiguous()
        call("xv2_loss_backward", ptr(logits), ptr(labels), n, h, w, ncls, lstride, post, terms, ptr(coef), ptr(up), ptr(dl))
        return dl, None, None, None, None, None


_LOSS_BITS = {"dice": lib.LOSS_DICE, "focal": lib.LOSS_FOCAL, "ce": lib.LOSS_CE, "ohem": lib.LOSS_CE}


def seg_loss(logits, labels, loss_str, post, weight=1.0, lstride=1):
    """Loss.forward (loss.py:85-101) for the dice / focal / ce / ohem terms.  'ce+ohem' counts CE twice like the reference."""
    if loss_str in ORDINAL_MODE:  # single-term ordinal heads (Loss.forward applies them alone, loss.py:92-101)
        return _OrdinalLoss.apply(logits, labels, ORDINAL_MODE[loss_str], post, weight, lstride)
    total = None
    fused = 0
    extra_ce = 0
    for name in loss_str.split("+"):
        if name not in _LOSS_BITS:
            raise NotImplementedError(f"loss '{name}' is outside the accelerated path (dice, focal, ce, ohem)")
        bit = _LOSS_BITS[name]
        if fused & bit:
            extra_ce += 1
        fused |= bit
    total = _SegLoss.apply(logits, labels, fused, post, weight, lstride)
    for _ in range(extra_ce):
        total = total + _SegLoss.apply(logits, labels, lib.LOSS_CE, post, weight, lstride)
    return total


class _OrdinalLoss(torch.autograd.Function):
    """'mse' (mode 0, loss.py:92-94) / 'coral' (mode 1, loss.py:54-65) with the `post` masking of loss.py:86-90."""

    @staticmethod
    def forward(ctx, logits, labels, mode, post, weight, lstride):
        _require_cuda(logits)
        logits = nhwc(logits.float())
        n, nl, h, w = logits.shape
        assert nl == (1 if mode == 0 else 3), "mse expects 1 logit per pixel, coral 3"
        if lstride != 1:
            labels = labels[:, ::lstride, ::lstride]
        labels = labels.contiguous()
        assert labels.dtype == torch.uint8 and labels.shape == (n, h, w)
        dev = logits.device
        sums = torch.zeros(2, dtype=torch.float64, device=dev)
        call("xv2_ordinal_loss_partials", ptr(logits), ptr(labels), n * h * w, mode, int(post), ptr(sums))
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        coef = torch.empty(1, dtype=torch.float32, device=dev)
        call("xv2_ordinal_loss_finalize", ptr(sums), float(weight), ptr(loss), ptr(coef))
        ctx.save_for_backward(logits, labels, coef)
        ctx.cfg = (mode, int(post))
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        logits, labels, coef = ctx.saved_tensors
        mode, post = ctx.cfg
        dl = torch.empty_like(logits)
        up = dloss.reshape(1).float().contiguous()
        call("xv2_ordinal_loss_backward", ptr(logits), ptr(labels), logits.shape[0] * logits.shape[2] * logits.shape[3], mode,
             post, ptr(coef), ptr(up), ptr(dl))
        return dl, None, None, None, None, None


ORDINAL_MODE = {"mse": 0, "coral": 1}


def ordinal_labels(logits, loss_str, labels=None, counters=None, clamp4=True, want_u8=False, want_f32=False):
    """convert_to_labels (utils/f1.py:7-15) for the mse / coral heads; optionally accumulates the F1 counters."""
    logits = nhwc(logits.detach().float())
    n, _, h, w = logits.shape
    u8 = torch.empty((n, h, w), dtype=torch.uint8, device=logits.device) if want_u8 else None
    f32 = torch.empty((n, h, w), dtype=torch.float32, device=logits.device) if want_f32 else None
    labels_c = labels.contiguous() if labels is not None else None
    call("xv2_ordinal_labels", ptr(logits), ptr(labels_c), n * h * w,
         ORDINAL_MODE[loss_str], int(clamp4), ptr(counters), ptr(u8), ptr(f32))
    return u8 if want_u8 else f32


def f1_update(logits, labels, n_class, counters, pred_map=None):
    """F1.update (utils/f1.py:28-42).  counters: int64 [3*(n_class-1)] = tp | fp | fn, accumulated in place."""
    logits = nhwc(logits.float())
    labels = labels.contiguous()
    n, _, h, w = logits.shape
    call("xv2_f1_update", ptr(logits), ptr(labels), n * h * w, n_class, ptr(counters), ptr(pred_map))


def mean4(a, b, c, d):
    a, b, c, d = (nhwc(t.float()) for t in (a, b, c, d))
    out = torch.empty_like(a)
    call("xv2_mean4", ptr(a), ptr(b), ptr(c), ptr(d), ptr(out), a.numel())
    return out


def post_process(loc_logits, dmg_logits):
    """Model.save + utils/post_process.py:27-38 from logits: returns (pre, post) uint8 maps (n, h, w)."""
    loc, dmg = nhwc(loc_logits.float()), nhwc(dmg_logits.float())
    n, _, h, w = loc.shape
    pre = torch.empty((n, h, w), dtype=torch.uint8, device=loc.device)
    post = torch.empty_like(pre)
    call("xv2_post_process", ptr(loc), ptr(dmg), n * h * w, ptr(pre), ptr(post))
    return pre, post


def post_process_probs(loc, dmg):
    """utils/post_process.py:27-38 from probabilities as the reference stores them: loc (h, w), dmg (4, h, w)."""
    loc, dmg = loc.contiguous().float(), dmg.contiguous().float()
    h, w = loc.shape
    pre = torch.empty((h, w), dtype=torch.uint8, device=loc.device)
    post = torch.empty_like(pre)
    call("xv2_post_process_probs", ptr(loc), ptr(dmg), h * w, ptr(pre), ptr(post))
    return pre, post


def cc_majority_vote(post_map):
    """utils/post_process.py:39-43 on the device: (n, h, w) | (h, w) uint8 damage map -> same shape, every 4-connected building
    carrying its majority class."""
    single = post_map.dim() == 2
    m = (post_map[None] if single else post_map).contiguous()
    n, h, w = m.shape
    lib.init(m.device.index)
    out = torch.empty_like(m)
    labels = torch.empty(n * h * w, dtype=torch.int32, device=m.device)
    votes = torch.empty((n * h * w, 4), dtype=torch.int32, device=m.device)
    call("xv2_cc_majority_vote", ptr(m), ptr(out), ptr(labels), ptr(votes), n, h, w)
    return out[0] if single else out


def dilate_square(label_map, k):
    """post_process.py:44-45 (skimage dilation(img, square(k))) on the device, k odd."""
    single = label_map.dim() == 2
    m = (label_map[None] if single else label_map).contiguous()
    n, h, w = m.shape
    lib.init(m.device.index)
    out = torch.empty_like(m)
    call("xv2_dilate_square", ptr(m), ptr(out), n, h, w, int(k))
    return out[0] if single else out


def score_counts(loc_pred, dmg_pred, loc_targ, dmg_targ, counters=None):
    """utils/xview2_metrics.py:61-92: accumulates the 15 TP / FN / FP counters of the xView2 scorer over uint8 label maps."""
    lib.init(loc_pred.device.index)
    if counters is None:
        counters = torch.zeros(15, dtype=torch.int64, device=loc_pred.device)
    lp, dp, lt, dt = (t.contiguous() for t in (loc_pred, dmg_pred, loc_targ, dmg_targ))  # named: copies outlive the launch call
    call("xv2_score_counts", ptr(lp), ptr(dp), ptr(lt), ptr(dt), lp.numel(), ptr(counters))
    return counters


def save_probs(logits):
    """Model.save (plt.py:126-131): (n, h, w) sigmoid(logit[:, 1]) for the 2-class head, (n, 4, h, w) softmax (plain
    NCHW, the layout np.save receives) for the 4-class head."""
    logits = nhwc(logits.detach().float())
    n, ncls, h, w = logits.shape
    shape = (n, h, w) if ncls == 2 else (n, ncls, h, w)
    out = torch.empty(shape, dtype=torch.float32, device=logits.device)
    call("xv2_save_probs", ptr(logits), n, h * w, ncls, ptr(out))
    return out


def normalize_tiles(pre_u8, post_u8=None, dtype=torch.bfloat16):
    """uint8 (n, h, w, 3) decoded tiles -> normalised (n, 3|6, h, w) channels-last activations (pytorch_loader.py:63,169-170)."""
    n, h, w, _ = pre_u8.shape
    lib.init(pre_u8.device.index)
    out = empty_act(n, 3 if post_u8 is None else 6, h, w, dtype, pre_u8.device)
    pre_c = pre_u8.contiguous()  # named, so that a temporary copy outlives the launch call (see _keep below)
    post_c = None if post_u8 is None else post_u8.contiguous()
    call("xv2_normalize_tiles", ptr(pre_c), ptr(post_c), ptr(out), n, h, w, dtype_code(out))
    return out


AUG_PARAMS = 16  # floats per sample, layout in include/xv2.h (xv2_augment_tiles)


def augment_tiles(pre_u8, post_u8, mask_u8, params, uniforms=None, crop=512, dtype=torch.bfloat16, want_u8=False):
    """Device-side train augmentation (pytorch_loader.py:73-92,124-148): uint8 (n, H, W, 3) decoded tiles [+ post] and the
    (n, H, W) mask -> normalised (n, 3|6, crop, crop) channels-last activations + the (n, crop, crop) mask crop.
    `params`: (n, 16) fp32 host-drawn decisions (see include/xv2.h); `uniforms` (n, 3): when given, the crop origin is chosen
    ON THE DEVICE from the mask content (CropNonEmptyMaskIfExists), else params[:, 4:6] is used."""
    n, sh, sw, _ = pre_u8.shape
    dev = pre_u8.device
    lib.init(dev.index)
    # every (possibly temporary) contiguous copy gets a NAME: a temporary created inside the call expression would be freed --
    # and its block handed to the next temporary -- before the kernel is even launched
    params = params.to(device=dev, dtype=torch.float32).contiguous()
    pre_c, mask_c = pre_u8.contiguous(), mask_u8.contiguous()
    post_c = None if post_u8 is None else post_u8.contiguous()
    origin = None
    if uniforms is not None:
        max_rows = int(sh * 1.3) + 2
        rowcount = torch.empty((n, max_rows), dtype=torch.int32, device=dev)
        origin = torch.empty((n, 2), dtype=torch.int32, device=dev)
        uni_c = uniforms.to(device=dev, dtype=torch.float32).contiguous()
        call("xv2_crop_origin", ptr(mask_c), ptr(params), ptr(uni_c), ptr(rowcount), ptr(origin), n, sh, sw, max_rows, crop, crop)
    ch = 3 if post_u8 is None else 6
    out = empty_act(n, ch, crop, crop, dtype, dev)
    out_u8 = torch.empty((n, crop, crop, ch), dtype=torch.uint8, device=dev) if want_u8 else None
    mask_out = torch.empty((n, crop, crop), dtype=torch.uint8, device=dev)
    call("xv2_augment_tiles", ptr(pre_c), ptr(post_c), ptr(mask_c), ptr(params), ptr(origin), ptr(out), ptr(out_u8),
         ptr(mask_out), n, sh, sw, crop, crop, dtype_code(out))
    return (out, mask_out, origin, out_u8) if want_u8 else (out, mask_out, origin)


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    call("xv2_adamw", ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), float(lr), float(beta1), float(beta2), float(eps),
         float(weight_decay), int(step), float(grad_scale))


def sgd_step(p, g, buf, lr, momentum, grad_scale, step):
    call("xv2_sgd", ptr(p), ptr(g), ptr(buf), p.numel(), float(lr), float(momentum), float(grad_scale), int(step))
