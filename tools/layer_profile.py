"""Per-launch table of one training step of the bench workload (CUDA events around every libxv2 call).

    python tools/layer_profile.py [--batch 8] [--size 1024] > gpurun_out/layers.txt
    python tools/layer_profile.py --ncu     # one step inside cudaProfilerStart/Stop for `ncu --profile-from-start off`
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from xview2_b200 import lib, ops
from xview2_b200.model.plt import Model

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--encoder", default="resnest50")
ap.add_argument("--ncu", action="store_true")
ap.add_argument("--fwd-only", action="store_true")
a = ap.parse_args()
a.gpus = 1
torch.cuda.set_device(0)
lib.init(0)
ns = bench.config_namespace(a)
torch.manual_seed(1)
model = Model(ns).cuda().train()
opt = model.configure_optimizers()
g = torch.Generator().manual_seed(1)
batch = {"tiles": torch.randint(0, 256, (a.batch, a.size, a.size, 3), generator=g, dtype=torch.uint8).cuda(),
         "mask": torch.randint(0, 2, (a.batch, a.size, a.size), generator=g, dtype=torch.uint8).cuda()}


def step():
    opt.zero_grad()
    loss = model.training_step(batch, 0)
    loss.backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
if a.ncu:
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
lib.profile_start()
step()
rows = lib.profile_stop(per_call=True)
total = sum(r[2] for r in rows)
print(f"# {len(rows)} launches, {total:.2f} ms")
print(f"{'idx':>5} {'entry':24s} {'ms':>8} {'TFLOP/s':>8} {'GB/s':>8}  tag")
for i, (name, tag, ms, fl, by) in enumerate(rows):
    tf = f"{fl / ms / 1e9:8.1f}" if fl and ms else " " * 8
    gb = f"{by / ms / 1e6:8.1f}" if by and ms else " " * 8
    print(f"{i:5d} {name:24s} {ms:8.4f} {tf} {gb}  {tag}")
