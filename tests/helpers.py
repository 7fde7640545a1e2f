"""Shared test helpers: golden fixtures, deterministic weights, digests."""
import argparse
import glob
import os

import torch

from oracle import functional as OF

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_DT = {"torch.float32": torch.float32, "torch.int64": torch.int64}


def golden_names():
    return sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.pt")))


def load_golden(name):
    fx = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)
    fx["ns"] = argparse.Namespace(**fx["args"])
    return fx


def golden_state(fx):
    """Seeded weights + the fixture's calibrated BN running statistics (stored under canonical keys)."""
    shapes = {k: (tuple(s), _DT[d]) for k, (s, d) in fx["state_shapes"].items()}
    state = OF.deterministic_state(shapes, fx["state_seed"])
    cal = fx.get("calibrated_running", {})
    for k in state:
        ck = OF.canonical_key(k)
        if ck in cal:
            state[k] = cal[ck].clone()
    return state


def golden_inputs(fx):
    g = torch.Generator().manual_seed(fx["input_seed"])
    ns, b, s = fx["ns"], fx["batch"], fx["size"]
    ch = 3 if ns.type == "pre" else 6
    x = torch.randn(b, ch, s, s, generator=g)
    hi = 2 if ns.type == "pre" else 5
    cells = torch.randint(0, hi, (b, s // 8, s // 8), generator=g, dtype=torch.uint8)  # same draw order as tools/make_golden.py
    y = cells.repeat_interleave(8, 1).repeat_interleave(8, 2).contiguous()
    return x, y


def check_digest(t, dg, rtol, atol_scale=1.0):
    """Compares tensor ``t`` with a stored digest (sum / norm / sampled entries)."""
    flat = t.detach().reshape(-1).double().cpu()
    norm = max(dg["norm"], 1e-12)
    assert abs(float(flat.norm()) - dg["norm"]) <= rtol * norm + 1e-7 * atol_scale, (float(flat.norm()), dg["norm"])
    ref = dg["val"].double()
    scale = max(float(ref.abs().max()), norm / max(flat.numel(), 1) ** 0.5, 1e-12)
    err = float((flat[dg["idx"]] - ref).abs().max())
    assert err <= rtol * scale * 4 + 1e-7 * atol_scale, (err, scale)


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def sub_logits(t, fx):
    """Applies the fixture's logit sub-sampling (--interpolate fixtures keep every 16th row / column) to a model output."""
    st = fx.get("logit_stride", 1)
    if st == 1:
        return t
    if isinstance(t, (list, tuple)):
        return [o[..., ::st, ::st] for o in t]
    return t[..., ::st, ::st]
