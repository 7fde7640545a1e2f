"""Real-PNG tile -> GPU loader throughput ("the real PNG loader measured separately").

Writes N synthetic 1024^2 satellite-like PNG tiles (pre / post + target masks) in the reference's directory layout, then times
    test  : TileLoader over TestDataset   (decode threads -> pinned ring -> side-stream H2D -> xv2_normalize_tiles)
    train : TileLoader over TrainPreDataset / TrainPostDataset with the device-side augmentation (xv2_augment_tiles)
and prints tiles/s with the number of decode threads used.   python tools/loader_bench.py [--tiles 64] [--workers N]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2
import numpy as np
import torch

from xview2_b200 import lib, ops
from xview2_b200.data_loading.pytorch_loader import fetch_pytorch_loader

ap = argparse.ArgumentParser()
ap.add_argument("--tiles", type=int, default=64)
ap.add_argument("--workers", type=int, default=os.cpu_count() or 8)
ap.add_argument("--batch", type=int, default=8)
a = ap.parse_args()
torch.cuda.set_device(0)
lib.init(0)
root = tempfile.mkdtemp(prefix="xv2_tiles_")
os.makedirs(os.path.join(root, "images"))
os.makedirs(os.path.join(root, "targets"))
rng = np.random.default_rng(0)
t0 = time.perf_counter()
png_bytes = 0
for i in range(a.tiles):
    for kind in ("pre", "post"):
        coarse = rng.integers(0, 256, (128, 128, 3)).astype(np.uint8)
        img = cv2.resize(coarse, (1024, 1024), interpolation=cv2.INTER_CUBIC).astype(np.int16) + rng.integers(-6, 7, (1024, 1024, 3))
        lbl = np.zeros((1024, 1024), np.uint8)
        for _ in range(12):
            y, x = rng.integers(0, 960, 2)
            lbl[y:y + rng.integers(10, 60), x:x + rng.integers(10, 60)] = rng.integers(1, 5) if kind == "post" else 1
        p = os.path.join(root, "images", f"tile{i:04d}_{kind}_disaster.png")
        cv2.imwrite(p, np.clip(img, 0, 255).astype(np.uint8))
        png_bytes += os.path.getsize(p)
        cv2.imwrite(os.path.join(root, "targets", f"tile{i:04d}_{kind}_disaster_target.png"), lbl)
print(f"wrote {a.tiles} pre/post pairs ({png_bytes / 2 / a.tiles / 1e6:.2f} MB per PNG) in {time.perf_counter() - t0:.1f} s", flush=True)


def run(mode, training):
    kw = dict(batch_size=a.batch, shuffle=training, drop_last=training, num_workers=a.workers, seed=1)
    loader = fetch_pytorch_loader(root, mode, training, kw)
    for b in loader:  # warm-up epoch (page cache, ring allocation)
        pass
    torch.cuda.synchronize()
    t = time.perf_counter()
    units = 0
    for b in loader:
        img = b["image"] if "image" in b else ops.normalize_tiles(b["tiles"], b.get("tiles_post"))
        units += img.shape[0]
    torch.cuda.synchronize()
    return units / (time.perf_counter() - t)


out = {"tiles": a.tiles, "decode_threads": a.workers, "host_cores": os.cpu_count(), "png_mb": round(png_bytes / 2 / a.tiles / 1e6, 2)}
out["test_pre_tiles_per_s"] = round(run("pre", False) * 1.0, 1)   # TestDataset decodes pre only in mode "pre"
out["test_post_pairs_per_s"] = round(run("post", False), 1)
out["train_pre_tiles_per_s_device_aug"] = round(run("pre", True), 1)
out["train_post_pairs_per_s_device_aug"] = round(run("post", True), 1)
os.environ["XV2_HOST_AUG"] = "1"
out["train_pre_tiles_per_s_host_aug"] = round(run("pre", True), 1)
print(json.dumps(out))
