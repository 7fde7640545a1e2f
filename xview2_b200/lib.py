"""ctypes binding of libxv2.so -- the C ABI declared in include/xv2.h.

The library is mandatory: importing this module (or calling any op) without the built extension raises.  There is no
CPU or PyTorch fallback behind these calls.
"""
import ctypes
import os
from ctypes import POINTER, c_double, c_float, c_int32, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxv2.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2
LOSS_DICE, LOSS_FOCAL, LOSS_CE = 1, 2, 4
E_UNSUPPORTED = -3


class ConvGeom(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in ("n", "h", "w", "c", "oh", "ow", "k", "r", "s", "stride", "pad", "dil", "ups",
                                        "groups", "dtype", "out_dtype")]


class TcConv(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in ("n", "h", "w", "c0", "c1", "ld0", "ld1", "k", "r", "s", "pad", "dil", "groups",
                                        "convt", "out_dtype", "ldo")]


P, I32, I64, F, D = c_void_p, c_int32, c_int64, c_float, c_double

_SIGNATURES = {
    "xv2_version": [],
    "xv2_init": [I32],
    "xv2_conv_gather_simt": [POINTER(ConvGeom), P, P, P, P, P],
    "xv2_conv_wgrad_simt": [POINTER(ConvGeom), P, P, P, P],
    "xv2_stem_conv_fwd": [P, P, P, I32, I32, I32, I32, P],
    "xv2_stem_conv_wgrad": [P, P, P, I32, I32, I32, I32, P],
    "xv2_fc_fwd": [P, P, P, P, I32, I32, I32, P],
    "xv2_fc_wgrad": [P, P, P, P, I32, I32, I32, P],
    "xv2_colsum": [P, I64, I32, I32, P, P],
    "xv2_pack_weight": [P, P, I32, I32, I32, I32, I32, I32, I32, P],
    "xv2_pack_weights_batched": [P, I32, P],
    "xv2_conv_tc": [POINTER(TcConv), P, P, P, P, P, P, P],
    "xv2_conv_tc_bnact": [POINTER(TcConv), P, P, P, P, P, I32, P, P],
    "xv2_wgrad_tc": [POINTER(TcConv), P, P, P, I32, P, P],
    "xv2_bn_stats": [P, I64, I32, I32, P, P],
    "xv2_bn_finalize": [P, I64, I32, P, P, P, P, F, F, P, P, P, P, P],
    "xv2_bn_eval_coeffs": [I32, P, P, P, P, F, P, P, P],
    "xv2_bn_apply": [P, P, P, I64, I32, I32, P, P, I32, P],
    "xv2_bn_train_apply": [P, P, P, I64, I32, I32, P, I64, P, P, P, P, F, F, P, I32, P],
    "xv2_bn_bwd_reduce": [P, P, P, I64, I32, I32, P, P, P, P, I32, P, P],
    "xv2_bn_bwd_reduce_du": [P, P, P, P, P, I64, I32, I32, P, P, P, P, I32, P, P],
    "xv2_bn_bwd_apply": [P, P, P, P, P, I64, I32, I32, P, P, P, P, P, I32, P, I64, P, P, I32, P],
    "xv2_maxpool_fwd": [P, P, P, I32, I32, I32, I32, I32, I32, I32, I32, I32, I32, P],
    "xv2_maxpool_bwd": [P, P, P, I32, I32, I32, I32, I32, I32, I32, I32, I32, I32, P],
    "xv2_avgpool_fwd": [P, P, I32, I32, I32, I32, I32, I32, I32, I32, I32, I32, I32, P],
    "xv2_avgpool_bwd": [P, P, I32, I32, I32, I32, I32, I32, I32, I32, I32, I32, I32, P],
    "xv2_splat_gap": [P, P, I32, I64, I32, I32, P],
    "xv2_rsoftmax_fwd": [P, P, I32, I32, P],
    "xv2_rsoftmax_bwd": [P, P, P, I32, I32, P],
    "xv2_splat_fc_fwd": [P, P, P, P, P, P, P, F, F, I32, P, P, P, P, P, P, I32, I32, I32, P],
    "xv2_splat_fc_bwd": [P, P, P, P, P, P, P, P, P, I32, P, P, P, P, P, P, P, P, P, I32, I32, I32, P],
    "xv2_splat_bn_gap_fin": [P, P, I64, P, P, P, P, F, F, P, P, P, I32, I64, I32, P],
    "xv2_splat_fc_bwd_fused": [P, P, P, P, P, P, I64, P, P, P, P, P, P, P, I32, P, P, P, P, P, P, P, P, P, P, I32, I32, I32, I32, P],
    "xv2_splat_combine": [P, P, P, I32, I64, I32, I32, P],
    "xv2_splat_bwd_att": [P, P, P, I32, I64, I32, I32, P],
    "xv2_splat_bwd_x": [P, P, P, P, I32, I64, I32, I32, P],
    "xv2_add_act": [P, P, P, I64, I32, I32, P],
    "xv2_act_bwd": [P, P, P, I64, I32, I32, P],
    "xv2_gate_fwd": [P, P, P, I64, I32, I32, P],
    "xv2_gate_bwd": [P, P, P, P, P, I64, I32, I32, P],
    "xv2_flip": [P, P, I32, I32, I32, I32, I32, I32, I32, P],
    "xv2_cast": [P, I32, P, I32, I64, P],
    "xv2_loss_partials": [P, P, I32, I32, I32, I32, I32, I32, P, P],
    "xv2_loss_finalize": [P, I32, I32, F, P, P, P],
    "xv2_loss_backward": [P, P, I32, I32, I32, I32, I32, I32, I32, P, P, P, P],
    "xv2_f1_update": [P, P, I64, I32, P, P, P],
    "xv2_mean4": [P, P, P, P, P, I64, P],
    "xv2_post_process": [P, P, I64, P, P, P],
    "xv2_post_process_probs": [P, P, I64, P, P, P],
    "xv2_save_probs": [P, I32, I64, I32, P, P],
    "xv2_cc_majority_vote": [P, P, P, P, I32, I32, I32, P],
    "xv2_dilate_square": [P, P, I32, I32, I32, I32, P],
    "xv2_score_counts": [P, P, P, P, I64, P, P],
    "xv2_splat_bn_gap": [P, P, P, P, I32, I64, I32, P],
    "xv2_splat_bn_combine": [P, P, P, P, P, I32, I64, I32, P],
    "xv2_splat_bn_bwd_partials": [P, P, P, P, P, I32, I64, I32, P],
    "xv2_splat_bn_bwd_datt": [P, P, P, P, I32, I32, P],
    "xv2_splat_bn_bwd_red": [P, P, P, P, P, P, I32, I64, I32, P],
    "xv2_splat_bn_bwd_apply": [P, P, P, P, P, P, P, P, P, P, P, P, P, I32, I32, I64, I32, P],
    "xv2_bnact_head_fwd": [P, I64, I32, P, P, I32, P, P, I32, P, P],
    "xv2_bnact_head_bwd_reduce": [P, P, I64, I32, P, P, P, P, I32, P, I32, P, P, P, P],
    "xv2_bnact_head_bwd_apply": [P, P, P, I64, I32, P, P, P, P, P, I32, P, I32, P, I64, P, P, I32, P],
    "xv2_adaptive_avgpool_fwd": [P, P, I32, I32, I32, I32, I32, I32, P],
    "xv2_adaptive_avgpool_bwd": [P, P, I32, I32, I32, I32, I32, I32, P],
    "xv2_bilinear_fwd": [P, P, I32, I32, I32, I32, I32, I32, I32, P],
    "xv2_bilinear_bwd": [P, P, I32, I32, I32, I32, I32, I32, I32, P],
    "xv2_ordinal_loss_partials": [P, P, I64, I32, I32, P, P],
    "xv2_ordinal_loss_finalize": [P, F, P, P, P],
    "xv2_ordinal_loss_backward": [P, P, I64, I32, I32, P, P, P, P],
    "xv2_ordinal_labels": [P, P, I64, I32, I32, P, P, P, P],
    "xv2_head_fwd": [P, P, P, P, I64, I32, I32, I32, P],
    "xv2_head_bwd": [P, P, P, P, P, P, I64, I32, I32, I32, P],
    "xv2_normalize_tiles": [P, P, P, I32, I32, I32, I32, P],
    "xv2_crop_origin": [P, P, P, P, P, I32, I32, I32, I32, I32, I32, P],
    "xv2_augment_tiles": [P, P, P, P, P, P, P, P, I32, I32, I32, I32, I32, I32, P],
    "xv2_adamw": [P, P, P, P, I64, F, F, F, F, F, I32, F, P],
    "xv2_sgd": [P, P, P, I64, F, F, F, I32, P],
    "xv2_adamw_dev": [P, P, P, P, I64, P, P],
    "xv2_sgd_dev": [P, P, P, I64, P, P],
}

_lib = None
_launches = 0  # kernels launched through this binding (bench.py reports it as gpu_launches)
_profile = None  # when a list: (name, start_event, end_event, flops, bytes) per call -- bench.py's roofline pass
_work = (0.0, 0.0, "")  # algorithmic (flops, bytes, tag) of the NEXT call, declared by ops.* through note_work()
_scope = ""  # profiling scope (e.g. "fwd:enc_l3") attached to every recorded call; set by scope_hooks() / set_scope()


class Xv2Error(RuntimeError):
    pass


def load():
    """Loads libxv2.so; raises if the extension has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Xv2Error(f"{LIB_PATH} is missing: build it with `python -m xview2_b200.build` "
                       "(the CUDA extension is mandatory, there is no CPU/PyTorch fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.xv2_last_error.restype = ctypes.c_char_p
    lib.xv2_last_error.argtypes = []
    for name, sig in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.restype = c_int32
        fn.argtypes = sig
    _lib = lib
    return lib


def exported_symbols():
    return ["xv2_last_error"] + list(_SIGNATURES)


def last_error():
    return load().xv2_last_error().decode()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device (raw C accessor: no Stream object per launch)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def dtype_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise Xv2Error(f"unsupported activation dtype {t.dtype}")


def note_work(flops=0.0, nbytes=0.0, tag=""):
    """Declares the algorithmic FLOPs / bytes (and a shape tag) of the next call (only read while profiling)."""
    global _work
    _work = (float(flops), float(nbytes), tag)


def profile_start():
    global _profile
    _profile = []


def set_scope(name):
    """Names the part of the step the following calls belong to (only read while profiling)."""
    global _scope
    _scope = name


def scope_hooks(named_modules, prefix="fwd:"):
    """Registers forward pre/post hooks that set the profiling scope to ``prefix + name`` while each module runs.
    Returns the hook handles (call .remove() on each when done)."""
    handles = []
    for name, mod in named_modules:
        def pre(_m, _i, name=name):
            set_scope(prefix + name)

        def post(_m, _i, _o):
            set_scope(prefix + "other")

        handles.append(mod.register_forward_pre_hook(pre))
        handles.append(mod.register_forward_hook(post))
    return handles


def profile_by_scope(rows):
    """{scope: {"ms", "flops", "bytes", "calls"}} from profile_stop(per_call=True, with_scope=True) rows."""
    out = {}
    for name, tag, ms, fl, by, scope in rows:
        d = out.setdefault(scope, {"calls": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        d["calls"] += 1
        d["ms"] += ms
        d["flops"] += fl
        d["bytes"] += by
    return out


def profile_stop(per_call=False, with_scope=False):
    """Returns {entry point: {"calls", "ms", "flops", "bytes"}} measured with CUDA events on the launching stream
    (per_call=True: the raw list of (entry point, tag, ms, flops, bytes) instead)."""
    global _profile
    rec, _profile = _profile, None
    torch.cuda.synchronize()
    if per_call and with_scope:
        return [(name, tag, e0.elapsed_time(e1), fl, by, sc) for name, e0, e1, fl, by, tag, sc in rec or []]
    if per_call:
        return [(name, tag, e0.elapsed_time(e1), fl, by) for name, e0, e1, fl, by, tag, _sc in rec or []]
    out = {}
    for name, e0, e1, fl, by, _tag, _sc in rec or []:
        d = out.setdefault(name, {"calls": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        d["calls"] += 1
        d["ms"] += e0.elapsed_time(e1)
        d["flops"] += fl
        d["bytes"] += by
    return out


def call(name, *args, allow_unsupported=False):
    """Calls an entry point with the current stream appended; raises Xv2Error on failure."""
    global _launches, _work
    lib = _lib if _lib is not None else load()
    if _profile is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args, stream_ptr())
        e1.record()
        if rc == 0:
            _profile.append((name, e0, e1, _work[0], _work[1], _work[2], _scope))
        _work = (0.0, 0.0, "")
    else:
        rc = getattr(lib, name)(*args, stream_ptr())
    if rc == 0:
        _launches += 1
        return 0
    if rc == E_UNSUPPORTED and allow_unsupported:
        return rc
    raise Xv2Error(f"{name} failed ({rc}): {lib.xv2_last_error().decode()}")


def launches():
    return _launches


def add_launches(n):
    """Accounts for kernels re-issued by a CUDA-graph replay (xview2_b200.graph) in the launch counter."""
    global _launches
    _launches += int(n)


_initialised = set()


def init(device=None):
    dev = torch.cuda.current_device() if device is None else int(device)
    if dev in _initialised:
        return
    lib = load()
    rc = lib.xv2_init(dev)
    if rc != 0:
        raise Xv2Error(f"xv2_init({dev}) failed ({rc}): {lib.xv2_last_error().decode()}")
    _initialised.add(dev)
