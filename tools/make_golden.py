"""Generates tests/golden/*.pt by running the REFERENCE'S OWN modules (model/unet.py, model/layers.py, model/loss.py,
imported verbatim from /root/reference through oracle/refload.py) on seeded inputs.  Build-container only: the GPU box
has no /root/reference, it consumes the committed fixtures.

    python tools/make_golden.py            # writes tests/golden/<case>.pt

Each fixture holds: the argparse fields, input/label seeds and shapes, the calibrated BN running statistics, eval-mode
logits, train-mode logits, the loss, per-parameter gradient digests (sum, L2 norm, 8 sampled entries) and the post-step
BN running statistics digests -- each from the reference modules in fp32 AND in fp64 (`*64`), plus `grad_noise`: the
reference's own fp32-vs-fp64 gradient error per parameter, the noise floor any fp32 implementation is held to.
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

# name: (arg overrides, batch, size).  Batch 4: ResNeSt's SplAt bn1 normalises a (B, C, 1, 1) tensor, and with B = 2
# x_hat is a sign function -- the fixture would be chaotic in fp32 (the reference's fp32 and fp64 runs differed by 3 %).
CASES = {
    "c1_resnet50_pre": (dict(encoder="resnet50"), 4, 64),
    "c2_resnest50_pre": (dict(encoder="resnest50"), 4, 128),
    "c2_resnest50_pre_ce_ohem": (dict(encoder="resnest50", loss_str="ce+ohem"), 4, 64),
    "c3_resnest50_siamese": (dict(encoder="resnest50", type="post", dmg_model="siamese"), 4, 64),
    "c3_resnest101_siamese": (dict(encoder="resnest101", type="post", dmg_model="siamese"), 4, 64),
    "c4_resnest50_fused_ds_attn": (dict(encoder="resnest50", type="post", dmg_model="fused", deep_supervision=True,
                                        attention=True), 4, 64),
    # BASELINE config 4's encoder at a size the CPU finishes: ResNeSt-200 fused + deep supervision + attention (386 M params)
    "c4_resnest200_fused_ds_attn": (dict(encoder="resnest200", type="post", dmg_model="fused", deep_supervision=True,
                                         attention=True, lite=True), 4, 64),
    # the remaining damage-model variants (unet.py:239-560), "lite" fixtures: logits + loss + sampled digests
    "v_resnest50_siameseEnc": (dict(type="post", dmg_model="siameseEnc", lite=True), 4, 64),
    "v_resnest50_fusedEnc": (dict(type="post", dmg_model="fusedEnc", lite=True), 4, 64),
    "v_resnest50_parallel": (dict(type="post", dmg_model="parallel", lite=True), 4, 64),
    "v_resnest50_parallelEnc": (dict(type="post", dmg_model="parallelEnc", lite=True), 4, 64),
    "v_resnet50_diff": (dict(encoder="resnet50", type="post", dmg_model="diff", lite=True), 4, 64),
    "v_resnest50_dilation2": (dict(dilation=2, lite=True), 4, 64),
    "v_resnest50_dilation4_noskip": (dict(dilation=4, no_skip=True, lite=True), 4, 64),
    "v_resnet50_dilation2": (dict(encoder="resnet50", dilation=2, lite=True), 4, 64),
    # optional model parts (layers.py:6-65,154,175-188; loss.py:54-65,92-94)
    "f4_resnet50_ppm": (dict(encoder="resnet50", ppm=True, lite=True), 4, 64),
    "f4_resnest50_aspp": (dict(aspp=True, lite=True), 4, 64),
    "f4_resnet50_dec_interp": (dict(encoder="resnet50", dec_interp=True, lite=True), 4, 64),
    # --interpolate resizes the logits to 512^2 (training) / 1024^2 (eval): the input must be a 512^2 crop for the training
    # loss to be defined; the fixture keeps every 16th logit row / column
    "f4_resnet50_interpolate": (dict(encoder="resnet50", interpolate=True, lite=True, logit_stride=16), 4, 512),
    "f4_resnest50_siamese_coral": (dict(type="post", dmg_model="siamese", loss_str="coral", lite=True), 4, 64),
    "f4_resnet50_siamese_mse": (dict(encoder="resnet50", type="post", dmg_model="siamese", loss_str="mse", lite=True), 4, 64),
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


def calibrate_running_stats(model, args, xc):
    """Gives the BN layers running statistics that match the activations they actually see (what a trained checkpoint
    has): ONE train-mode pass of the reference model with momentum None (cumulative average: a Siamese net that runs its
    shared U-Net on the pre and then the post image ends up with the mean of both) on the fixture's input.  With arbitrary
    running statistics the eval-mode residual stream grows ~2x per block (logits ~1e4) and the split-attention
    soft-max saturates, which makes the fixture chaotic: the reference's own fp32 and fp64 runs then differ by 19 %."""
    bns = [m for m in model.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    old = [m.momentum for m in bns]
    for m in bns:
        m.momentum = None
        m.num_batches_tracked.zero_()
    model.train()
    with torch.no_grad():
        model(xc)
    for m, mom in zip(bns, old):
        m.momentum = mom
        m.num_batches_tracked.zero_()
    return {k: v.clone() for k, v in model.state_dict().items() if "running_" in k and OF.canonical_key(k) == k}


def forward_backward(model, loss_mod, args, x, y):
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
    elif args.loss_str == "mse" and out.dtype == torch.float64:
        # fp64 yard-stick only: Loss.forward casts the target with .float() (loss.py:94), which nn.MSELoss rejects against
        # fp64 predictions; the same formula (loss.py:86-94) with a double target
        keep = y > 0
        pred = torch.relu(torch.stack([out[:, i][keep] for i in range(out.shape[1])], 1)[:, 0])
        loss = torch.nn.functional.mse_loss(pred, (y[keep] - 1).double())
        train_logits = out.detach().clone()
    else:
        loss = loss_mod(out, y)
        train_logits = out.detach().clone()
    loss.backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in model.named_parameters()}
    return train_logits, float(loss), grads


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def run_case(ref_unet, ref_loss, name, over, batch, size):
    import copy
    over = dict(over)
    lite = over.pop("lite", False)
    lstride = over.pop("logit_stride", 1)
    args = argparse.Namespace(**{**BASE, **over})
    model = build_reference_model(ref_unet, args)
    shapes = {k: (tuple(v.shape), v.dtype) for k, v in model.state_dict().items()}
    state = OF.deterministic_state(shapes)
    model.load_state_dict(state, strict=True)
    x, y = make_inputs(args, batch, size)
    running = calibrate_running_stats(model, args, x)
    loss_mod = ref_loss.Loss(args)
    model64 = copy.deepcopy(model).double()  # the reference modules in fp64: the conditioning yard-stick

    model.eval()
    model64.eval()
    with torch.no_grad():
        eval_logits = model(x).clone()
        eval_logits64 = model64(x.double()).clone()

    train_logits, loss, grads = forward_backward(model, loss_mod, args, x, y)
    train_logits64, loss64, grads64 = forward_backward(model64, loss_mod, args, x.double(), y)
    # named_parameters de-duplicates shared modules (FusedUNet)
    uniq = {k: (digest(g) if g is not None else None) for k, g in grads.items()}
    uniq64 = {k: (digest(g) if g is not None else None) for k, g in grads64.items()}
    noise = {k: (rel(grads[k], grads64[k]) if grads[k] is not None else None) for k in grads}
    post_state = {k: digest(v.float()) for k, v in model.state_dict().items() if "running_" in k}
    if lite:  # every 7th parameter / running statistic: keeps the many variant fixtures small
        keep = set(sorted(uniq)[::7])
        uniq = {k: v for k, v in uniq.items() if k in keep}
        uniq64 = {k: v for k, v in uniq64.items() if k in keep}
        noise = {k: v for k, v in noise.items() if k in keep}
        post_state = {k: v for k, v in post_state.items() if k in set(sorted(post_state)[::7])}
    as32 = lambda t: [o.float() for o in t] if isinstance(t, list) else t.float()
    l32 = train_logits[0] if isinstance(train_logits, list) else train_logits
    l64 = train_logits64[0] if isinstance(train_logits64, list) else train_logits64
    print(f"  conditioning (reference fp32 vs fp64): eval logits {rel(eval_logits, eval_logits64):.2e}  train logits "
          f"{rel(l32, l64):.2e}  grads median {sorted(v for v in noise.values() if v is not None)[len(noise) // 2]:.2e}")
    if lstride > 1:
        sub = lambda t: [o[..., ::lstride, ::lstride].clone() for o in t] if isinstance(t, list) else t[..., ::lstride, ::lstride].clone()
        eval_logits, eval_logits64, train_logits, train_logits64 = (sub(t) for t in (eval_logits, eval_logits64, train_logits, train_logits64))
    return {
        "name": name, "args": vars(args), "batch": batch, "size": size, "input_seed": 1, "state_seed": 1, "logit_stride": lstride,
        "state_shapes": {k: (s, str(d)) for k, (s, d) in shapes.items()},
        "calibrated_running": running,
        "eval_logits": eval_logits, "eval_logits64": eval_logits64.float(),
        "train_logits": train_logits, "train_logits64": as32(train_logits64), "loss": loss, "loss64": loss64,
        "grad_digest": uniq, "grad_digest64": uniq64, "grad_noise": noise, "running_digest": post_state,
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
