"""Debug aid (GPU box): gradient w.r.t. every stage output, CUDA fp32 path vs CPU oracle (train mode, 'pre' models)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from oracle import functional as OF
from tests.helpers import golden_inputs, golden_state, load_golden, rel_err

torch.backends.cudnn.allow_tf32 = False
name = sys.argv[1] if len(sys.argv) > 1 else "c1_resnet50_pre"
fx = load_golden(name)
ns = fx["ns"]
ns.precision = 32
from xview2_b200.model.unet import UNetLoc
from xview2_b200.model.loss import Loss
from xview2_b200 import ops
from xview2_b200.lib import ACT_LRELU

model = UNetLoc(ns)
state = golden_state(fx)
model.load_state_dict(state, strict=True)
model = model.cuda().train()
x, y = golden_inputs(fx)
stages = ["enc_l1", "enc_l2", "enc_l3", "enc_l4", "enc_l5", "dec_l1", "dec_l2", "dec_l3", "dec_l4", "dec_l5"]
grads, acts = {}, {}
hooks = []
for nm in stages:
    def fh(m, i, o, nm=nm):
        acts[nm] = o.detach().float().cpu()
        o.register_hook(lambda g, nm=nm: grads.__setitem__(nm, g.detach().float().cpu()))
    hooks.append(getattr(model.unet, nm).register_forward_hook(fh))
out = model(x.cuda())
out.register_hook(lambda g: grads.__setitem__("logits", g.detach().float().cpu()))
loss = Loss(ns)(out, y.cuda())
loss.backward()

P = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k, v in state.items()}
encs = OF.encoder_forward(P, "unet.", x, True, ns.encoder, 1)
names = [f"unet.dec_l{i}" for i in range(1, 6)]
d1 = OF.upsample_block(P, names[0], encs[4], encs[3], True, ns.attention)
d2 = OF.upsample_block(P, names[1], d1, encs[2], True, ns.attention)
d3 = OF.upsample_block(P, names[2], d2, encs[1], True, ns.attention)
d4 = OF.upsample_block(P, names[3], d3, encs[0], True, ns.attention)
d5 = OF.upsample_block(P, names[4], d4, None, True, ns.attention)
ref = dict(zip(stages, encs + [d1, d2, d3, d4, d5]))
for t in ref.values():
    t.retain_grad()
logits = OF.output_template(P, "output_block", d5, d4, d3, True, False)
logits.retain_grad()
rl = OF.compute_loss(logits, y, ns.loss_str, False, False)
rl.backward()
print("loss", float(loss.detach()), float(rl.detach()))
print("dlogits", rel_err(grads["logits"], logits.grad))
for nm in reversed(stages):
    print(f"{nm:8s} act err {rel_err(acts[nm], ref[nm]):.3e}  grad err {rel_err(grads[nm], ref[nm].grad):.3e}  grad absmax {float(ref[nm].grad.abs().max()):.3e}")

# isolated ConvLayer backward at dec_l5 scale with the real tensors
blk = model.unet.dec_l5.conv_block.conv2
xin = torch.randn(2, 32, 64, 64)
gy = torch.randn(2, 32, 64, 64)
xi = ops.nhwc(xin.cuda()).requires_grad_(True)
for p in blk.parameters():
    p.grad = None
o = blk(xi)
o.backward(ops.nhwc(gy.cuda()))
k = "unet.dec_l5.conv_block.conv2"
P2 = {kk: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in kk else v.clone()) for kk, v in state.items() if kk.startswith(k)}
xr = xin.clone().requires_grad_(True)
orf = OF.conv_layer(P2, k, xr, True)
orf.backward(gy)
print("isolated ConvLayer: fwd", rel_err(o, orf), "dx", rel_err(xi.grad, xr.grad), "dw", rel_err(blk.conv.weight.grad, P2[k + ".conv.weight"].grad),
      "dgamma", rel_err(blk.batch_norm.weight.grad, P2[k + ".batch_norm.weight"].grad), "dbeta", rel_err(blk.batch_norm.bias.grad, P2[k + ".batch_norm.bias"].grad))
