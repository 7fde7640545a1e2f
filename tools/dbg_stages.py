"""Debug aid (GPU box): per-stage activation error and per-parameter gradient error of the fp32 CUDA path against the
CPU oracle for one golden fixture.   python tools/dbg_stages.py <fixture> [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import functional as OF
from tests.helpers import golden_inputs, golden_state, load_golden, rel_err

torch.backends.cudnn.allow_tf32 = False
name = sys.argv[1] if len(sys.argv) > 1 else "c2_resnest50_pre"
prec = sys.argv[2] if len(sys.argv) > 2 else "32"
fx = load_golden(name)
ns = fx["ns"]
ns.precision = 32 if prec == "32" else prec
from xview2_b200.model.unet import UNetLoc, get_dmg_unet
from xview2_b200.model.loss import Loss
from xview2_b200.model.plt import compute_loss

model = UNetLoc(ns) if ns.type == "pre" else get_dmg_unet(ns)
state = golden_state(fx)
model.load_state_dict(state, strict=True)
model = model.cuda()
x, y = golden_inputs(fx)

if ns.type == "pre":
    for training in (False, True):
        model.train(training)
        caps = {}
        hooks = []
        for nm in ["enc_l1", "enc_l2", "enc_l3", "enc_l4", "enc_l5", "dec_l1", "dec_l2", "dec_l3", "dec_l4", "dec_l5"]:
            hooks.append(getattr(model.unet, nm).register_forward_hook(lambda m, i, o, nm=nm: caps.__setitem__(nm, o.detach().float().cpu())))
        sd0 = {k: v.clone() for k, v in model.state_dict().items()}
        with torch.no_grad():
            out = model(x.cuda())
        model.load_state_dict(sd0)
        for h in hooks:
            h.remove()
        P = {k: v.clone() for k, v in state.items()}
        with torch.no_grad():
            encs = OF.encoder_forward(P, "unet.", x, training, ns.encoder, 1)
            names = [f"unet.dec_l{i}" for i in range(1, 6)]
            d1 = OF.upsample_block(P, names[0], encs[4], encs[3], training, ns.attention)
            d2 = OF.upsample_block(P, names[1], d1, encs[2], training, ns.attention)
            d3 = OF.upsample_block(P, names[2], d2, encs[1], training, ns.attention)
            d4 = OF.upsample_block(P, names[3], d3, encs[0], training, ns.attention)
            d5 = OF.upsample_block(P, names[4], d4, None, training, ns.attention)
        refs = dict(zip(["enc_l1", "enc_l2", "enc_l3", "enc_l4", "enc_l5", "dec_l1", "dec_l2", "dec_l3", "dec_l4", "dec_l5"],
                        encs + [d1, d2, d3, d4, d5]))
        for nm, r in refs.items():
            print(f"training={training} {nm:8s} shape {tuple(r.shape)} rel err {rel_err(caps[nm], r):.3e}")

# gradients
model.train()
model.load_state_dict(state)
out = model(x.cuda())
loss = compute_loss(Loss(ns), out, y.cuda(), ns.deep_supervision)
loss.backward()
P = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k, v in state.items()}
ref_out = OF.model_forward(P, x, True, ns)
ref_loss = OF.compute_loss(ref_out, y, ns.loss_str, ns.type == "post", ns.deep_supervision)
ref_loss.backward()
print("loss", float(loss), float(ref_loss))
worst = []
for k, p in model.named_parameters():
    if OF.canonical_key(k) != k:
        continue
    g = P[k].grad
    if g is None or p.grad is None:
        print("missing grad", k, g is None, p.grad is None)
        continue
    worst.append((rel_err(p.grad, g), k, float(g.abs().max())))
for e, k, m in worst:
    flag = " <<<" if e > 2e-3 else ""
    print(f"grad {k:60s} rel {e:.3e} max {m:.3e}{flag}")
