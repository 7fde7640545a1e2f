"""Inference throughput of BASELINE config 5 (ResNeSt-50 U-Net, eval, optional --tta x4, argmax label map + F1 counters) on one
GPU: tiles/s with inputs resident in HBM.   python tools/eval_bench.py [--batch 16] [--tta] [--steps 10]"""
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
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--encoder", default="resnest50")
ap.add_argument("--tta", action="store_true")
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
a.gpus = 1
torch.cuda.set_device(0)
lib.init(0)
ns = bench.config_namespace(a)
ns.tta = a.tta
torch.manual_seed(1)
model = Model(ns).cuda().eval()
g = torch.Generator().manual_seed(1)
batch = {"tiles": torch.randint(0, 256, (a.batch, a.size, a.size, 3), generator=g, dtype=torch.uint8).cuda(),
         "mask": torch.randint(0, 2, (a.batch, a.size, a.size), generator=g, dtype=torch.uint8).cuda()}
pred_map = torch.empty((a.batch, a.size, a.size), dtype=torch.uint8, device="cuda")


def step():
    with torch.no_grad():
        img = model._image(batch)
        pred = model.forward(img)
        model.f1_score.update(pred, batch["mask"], pred_map)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n0 = lib.launches()
e0.record()
for _ in range(a.steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
passes = 4 if a.tta else 1
print(f"eval resnest50 batch {a.batch} tta={a.tta}: {ms:.2f} ms/step, {a.batch / ms * 1e3:.1f} tiles/s, "
      f"{(lib.launches() - n0) // a.steps} launches/step, "
      f"{578.8 * passes * a.batch / ms:.0f} TFLOP/s = {578.8 * passes * a.batch / ms / bench.measured_peaks()['tensor']:.3f} of the sustained bf16 peak")
