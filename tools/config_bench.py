"""One-GPU training-step timing of any BASELINE.json configuration at full size (synthetic 1024^2 tiles / pairs resident in HBM).

    python tools/config_bench.py --name C3 --type post --dmg_model siamese --encoder resnest101 --batch 4
    python tools/config_bench.py --name C4 --type post --dmg_model fused --encoder resnest200 --batch 2 --deep_supervision --attention
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from xview2_b200 import lib
from xview2_b200.model.plt import Model

ap = argparse.ArgumentParser()
ap.add_argument("--name", default="cfg")
ap.add_argument("--type", default="pre")
ap.add_argument("--dmg_model", default="siamese")
ap.add_argument("--encoder", default="resnest50")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--deep_supervision", action="store_true")
ap.add_argument("--attention", action="store_true")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--graph", action="store_true", help="capture forward+backward in a CUDA graph")
ap.add_argument("--gflop", type=float, default=0.0, help="fwd+bwd GFLOP per unit (SURVEY.md 8d) for the roofline fraction")
a = ap.parse_args()
a.gpus = 1
torch.cuda.set_device(0)
lib.init(0)
ns = bench.config_namespace(a)
ns.type, ns.dmg_model, ns.deep_supervision, ns.attention = a.type, a.dmg_model, a.deep_supervision, a.attention
torch.manual_seed(1)
model = Model(ns).cuda().train()
opt = model.configure_optimizers()
g = torch.Generator().manual_seed(1)
post = a.type == "post"
batch = {"tiles": torch.randint(0, 256, (a.batch, a.size, a.size, 3), generator=g, dtype=torch.uint8).cuda(),
         "mask": torch.randint(0, 5 if post else 2, (a.batch, a.size, a.size), generator=g, dtype=torch.uint8).cuda()}
if post:
    batch["tiles_post"] = torch.randint(0, 256, (a.batch, a.size, a.size, 3), generator=g, dtype=torch.uint8).cuda()


def step():
    opt.zero_grad()
    loss = model.training_step(batch, 0)
    loss.backward()
    opt.step()
    return loss


if a.graph:
    from xview2_b200.graph import GraphedTrainStep
    gstep = GraphedTrainStep(model, opt, batch)
    step = lambda: gstep(batch)  # noqa: E731
for _ in range(3):
    step()
torch.cuda.synchronize()
n0 = lib.launches()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
params = sum(p.numel() for p in model.parameters()) / 1e6
frac = f", {a.gflop * a.batch / ms:.0f} TFLOP/s = {a.gflop * a.batch / ms / bench.measured_peaks()['tensor']:.3f} of the sustained bf16 peak" if a.gflop else ""
print(f"{a.name}{" [graph]" if a.graph else ""}: {a.encoder} {a.type}/{a.dmg_model} ds={a.deep_supervision} attn={a.attention} batch {a.batch}: {ms:.1f} ms/step, "
      f"{a.batch / ms * 1e3:.1f} {'pairs' if post else 'tiles'}/s, {(lib.launches() - n0) // a.steps} launches/step, {params:.1f} M params, "
      f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, loss {float(loss):.4f}{frac}")
