"""bf16 train-mode forward error of the smoke configuration vs the fp32 oracle, beside PyTorch's own bf16 autocast."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
from oracle import functional as OF
from xview2_b200 import lib, ops
from xview2_b200.model.unet import UNetLoc

torch.cuda.set_device(0); lib.init(0)
for size, batch in ((64, 4), (128, 4), (64, 8)):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(batch, 3, size, size, generator=g)
    args = ge._args(precision="bf16")
    model = UNetLoc(args)
    shapes = {k: (tuple(v.shape), v.dtype) for k, v in model.state_dict().items()}
    state = OF.deterministic_state(shapes, 1)
    P = {k: v.clone() for k, v in state.items()}
    ref = OF.model_forward(P, x, True, args)
    P64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in state.items()}
    ref64 = OF.model_forward(P64, x.double(), True, args).float()
    model.load_state_dict(state, strict=True)
    model = model.cuda().train()
    with torch.no_grad():
        out = model(x.cuda()).float().cpu()
    Pg = {k: v.cuda() for k, v in state.items()}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        amp = OF.model_forward(Pg, x.cuda(), True, args).float().cpu()
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print(f"size {size} batch {batch}: ours vs fp64 {rel(out, ref64):.4f}  torch-autocast vs fp64 {rel(amp, ref64):.4f}  fp32 vs fp64 {rel(ref, ref64):.5f}")
