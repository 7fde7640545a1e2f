"""Generates tests/golden/*.pt by running the REFERENCE'S OWN modules (model/unet.py, model/layers.py, model/loss.py,
imported verbatim from /root/reference through oracle/refload.py) on seeded inputs.  Build-container only: the GPU box
has no /root/reference, it consumes the committed fixtures.

    python tools/make_golden.py            # writes tests/golden/<case>.pt

Each fixture holds: the argparse fields, input/label seeds and shapes, eval-mode logits, train-mode logits, the loss,
per-parameter gradient digests (sum, L2 norm, 8 sampled entries) and the post-step BN running statistics digests.
Weights are NOT stored: both sides regenerate them with oracle.functional.deterministic_state (crc32-seeded per key).
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import functional as OF  # noqa: E402
from oracle.refload import load_reference  # noqa: E402

BASE = dict(ppm=False, aspp=False, dilation=1, no_skip=False, interpolate=False, attention=False, dec_interp=False,
            deep_supervision=False, loss_str="focal+dice", encoder="resnest50", dmg_model="siamese", type="pre", tta=False)

CASES = {
    # name: (arg overrides, batch, size)
    "c1_resnet50_pre": (dict(encoder="resnet50"), 2, 64),
    "c2_resnest50_pre": (dict(encoder="resnest50"), 2, 128),
    "c2_resnest50_pre_ce_ohem": (dict(encoder="resnest50", loss_str="ce+ohem"), 2, 64),
    "c3_resnest50_siamese": (dict(encoder="resnest50", type="post", dmg_model="siamese"), 2, 64),
    "c3_resnest101_siamese": (dict(encoder="resnest101", type="post", dmg_model="siamese"), 2, 64),
    "c4_resnest50_fused_ds_attn": (dict(encoder="resnest50", type="post", dmg_model="fused", deep_supervision=True,
                                        attention=True), 2, 64),
}


def make_inputs(args, batch, size, seed=1):
    g = torch.Generator().manual_seed(seed)
    ch = 3 if args.type == "pre" else 6
    x = torch.randn(batch, ch, size, size, generator=g)
    hi = 2 if args.type == "pre" else 5
    # blocky labels (nearest-upsampled 8x8 cells) so classes form regions like building footprints
    cells = torch.randint(0, hi, (batch, size // 8, size // 8), generator=g, dtype=torch.uint8)
    y = cells.repeat_interleave(8, 1).repeat_interleave(8, 2).contiguous()
    return x, y


def digest(t, n=8):
    flat = t.detach().reshape(-1).double()
    g = torch.Generator().manual_seed(flat.numel())
    idx = torch.randint(0, flat.numel(), (min(n, flat.numel()),), generator=g)
    return {"sum": float(flat.sum()), "norm": float(flat.norm()), "idx": idx, "val": flat[idx].float().clone()}


def build_reference_model(ref_unet, args):
    torch.manual_seed(1)
    return ref_unet.UNetLoc(args) if args.type == "pre" else ref_unet.get_dmg_unet(args)


def run_case(ref_unet, ref_loss, name, over, batch, size):
    args = argparse.Namespace(**{**BASE, **over})
    model = build_reference_model(ref_unet, args)
    shapes = {k: (tuple(v.shape), v.dtype) for k, v in model.state_dict().items()}
    state = OF.deterministic_state(shapes)
    model.load_state_dict(state, strict=True)
    x, y = make_inputs(args, batch, size)
    loss_mod = ref_loss.Loss(args)

    model.eval()
    with torch.no_grad():
        eval_logits = model(x).clone()

    model.train()
    out = model(x)
    # Model.compute_loss, plt.py:69-77 (plt.py itself cannot be imported here: needs pytorch_lightning/apex)
    if args.deep_supervision:
        loss = loss_mod(out[0], y)
        for i, pred in enumerate(out[1:]):
            ds = torch.nn.functional.interpolate(y.unsqueeze(1), pred.shape[2:])
            loss = loss + 0.5 ** (i + 1) * loss_mod(pred, ds.squeeze(1))
        loss = loss / (2 - 2 ** (-len(out)))
        train_logits = [o.detach().clone() for o in out]
    else:
        loss = loss_mod(out, y)
        train_logits = out.detach().clone()
    loss.backward()
    uniq = {}
    for k, p in model.named_parameters():  # named_parameters de-duplicates shared modules (FusedUNet)
        uniq[k] = digest(p.grad) if p.grad is not None else None
    post_state = {k: digest(v.float()) for k, v in model.state_dict().items() if "running_" in k}
    return {
        "name": name, "args": vars(args), "batch": batch, "size": size, "input_seed": 1, "state_seed": 1,
        "state_shapes": {k: (s, str(d)) for k, (s, d) in shapes.items()},
        "eval_logits": eval_logits, "train_logits": train_logits, "loss": float(loss),
        "grad_digest": uniq, "running_digest": post_state,
        "torch": torch.__version__,
    }


def main():
    ref_unet, _, ref_loss = load_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = sys.argv[1:]
    for name, (over, batch, size) in CASES.items():
        if only and name not in only:
            continue
        fx = run_case(ref_unet, ref_loss, name, over, batch, size)
        path = os.path.join(out_dir, name + ".pt")
        torch.save(fx, path)
        print(f"{name}: loss={fx['loss']:.6f} params={len(fx['grad_digest'])} -> {os.path.getsize(path)/1e6:.2f} MB")


if __name__ == "__main__":
    main()
