"""GPU tests (-m gpu) of the device-side train augmentation (xview2_b200/csrc/augment.cu) against the float32
per-output-pixel restatement in oracle/functional.py (itself pinned to cv2.resize in tests/test_host.py)."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import functional as OF

pytestmark = pytest.mark.gpu


def _tiles(n, h, w, seed):
    import cv2
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        coarse = rng.integers(0, 256, (h // 8, w // 8, 3)).astype(np.uint8)
        img = cv2.resize(coarse, (w, h), interpolation=cv2.INTER_LINEAR).astype(np.int16) + rng.integers(-12, 13, (h, w, 3))
        out.append(np.clip(img, 0, 255).astype(np.uint8))
    return np.stack(out)


def _params(n, h, w, seed, noise=False):
    from xview2_b200.data_loading.pytorch_loader import GpuTrainAugment
    aug = GpuTrainAugment(128)
    rng = random.Random(seed)
    P = np.stack([aug.draw(rng, h, w, 2) for _ in range(n)])
    P[0::2, 15] = 1.0  # force zoom on every other sample
    for i in range(0, n, 2):
        s = 1.0 + 0.3 * rng.random()
        P[i, 2], P[i, 3] = int(w * s), int(h * s)
        P[i, 0], P[i, 1] = w / P[i, 2], h / P[i, 3]
    P[:, 10:14] = np.array([[1.1, -0.07, 0.85, 0.12]], np.float32)  # brightness / contrast on for both images
    P[:, 8:10] = np.array([[4.0, 6.5]], np.float32) if noise else 0.0
    return P.astype(np.float32)


@pytest.mark.parametrize("with_post", [False, True])
def test_augment_tiles_bit_exact_without_noise(with_post):
    from xview2_b200 import ops
    n, h, w, crop = 6, 160, 192, 128
    pre, post = _tiles(n, h, w, 1), (_tiles(n, h, w, 2) if with_post else None)
    rng = np.random.default_rng(3)
    mask = (rng.random((n, h, w)) > 0.98).astype(np.uint8) * rng.integers(1, 5, (n, h, w)).astype(np.uint8)
    mask[3] = 0  # an empty mask: uniform crop origin
    P = _params(n, h, w, 5)
    out, mask_out, origin, out_u8 = ops.augment_tiles(torch.from_numpy(pre).cuda(), None if post is None else torch.from_numpy(post).cuda(),
                                                      torch.from_numpy(mask).cuda(), torch.from_numpy(P[:, :16]).cuda(),
                                                      torch.from_numpy(P[:, 16:19]).cuda(), crop=crop, dtype=torch.float32, want_u8=True)
    origin = origin.cpu().numpy()
    for i in range(n):
        want_origin = OF.crop_origin_restatement(mask[i], P[i], P[i, 16:19], crop)
        assert tuple(origin[i]) == want_origin, (i, origin[i], want_origin)
        u8, norm, m = OF.augment_restatement(pre[i], None if post is None else post[i], mask[i], P[i], origin[i], crop)
        assert np.array_equal(out_u8[i].cpu().numpy(), u8), f"sample {i}: augmented bytes differ"
        assert np.array_equal(mask_out[i].cpu().numpy(), m)
        got = out[i].permute(1, 2, 0).cpu().numpy()
        assert np.array_equal(got, norm.astype(np.float32))
        if mask[i].any():
            assert m.any(), "CropNonEmptyMaskIfExists: the crop must contain a building pixel"


def test_augment_tiles_noise_statistics_and_restatement():
    from xview2_b200 import ops
    n, h, w, crop = 4, 160, 160, 128
    pre, post = _tiles(n, h, w, 7), _tiles(n, h, w, 8)
    mask = np.ones((n, h, w), np.uint8)
    P = _params(n, h, w, 9, noise=True)
    P[:, 15] = 0
    P[:, :4] = [1, 1, w, h]
    P[:, 10:14] = [1, 0, 1, 0]
    _, _, origin, out_u8 = ops.augment_tiles(torch.from_numpy(pre).cuda(), torch.from_numpy(post).cuda(), torch.from_numpy(mask).cuda(),
                                             torch.from_numpy(P[:, :16]).cuda(), torch.from_numpy(P[:, 16:19]).cuda(), crop=crop,
                                             want_u8=True)
    origin = origin.cpu().numpy()
    got = out_u8.cpu().numpy().astype(np.int32)
    bad = 0
    for i in range(n):
        OF.augment_restatement.sample_base = i * crop * crop
        u8, _, _ = OF.augment_restatement(pre[i], post[i], mask[i], P[i], origin[i], crop)
        d = np.abs(got[i] - u8.astype(np.int32))
        assert d.max() <= 1  # libm vs CUDA log / cos differ in the last bits: at most one grey level, rarely
        bad += int((d > 0).sum())
    OF.augment_restatement.sample_base = 0
    assert bad / got.size < 2e-3
    # statistics of the added noise on pixels away from saturation: zero mean, variance sigma^2 (+ 1/12 from truncation)
    for im, sigma in ((0, 4.0), (1, 6.5)):
        diffs = []
        for i in range(n):
            x0, y0 = origin[i]
            src = (pre if im == 0 else post)[i]
            view = src[y0:y0 + crop, x0:x0 + crop].astype(np.int32)
            if P[i, 7]:
                view = view[::-1]
            if P[i, 6]:
                view = view[:, ::-1]
            ok = (view > 40) & (view < 215)
            diffs.append((got[i][:, :, im * 3:im * 3 + 3] - view)[ok])
        d = np.concatenate(diffs).astype(np.float64)
        assert abs(d.mean() + 0.5) < 0.1          # truncation toward zero of a positive value: -0.5 on average
        assert abs(d.var() - (sigma ** 2 + 1 / 12)) < 0.06 * sigma ** 2


def test_tile_loader_augments_on_device(tmp_path):
    """TileLoader with the default (device) augmentation: decode threads -> pinned ring (full tiles + 19 floats) -> H2D ->
    one gather kernel -> {"image": normalised 512^2 crops, "mask"}."""
    import cv2

    from xview2_b200.data_loading.pytorch_loader import fetch_pytorch_loader
    rng = np.random.default_rng(0)
    os.makedirs(tmp_path / "images")
    os.makedirs(tmp_path / "targets")
    for i in range(5):
        for kind in ("pre", "post"):
            img = rng.integers(0, 256, (1024, 1024, 3), dtype=np.uint8)
            lbl = np.zeros((1024, 1024), np.uint8)
            lbl[100 + 150 * i:160 + 150 * i, 700:800] = 1 + (i % 4 if kind == "post" else 0)
            cv2.imwrite(str(tmp_path / "images" / f"t{i}_{kind}_disaster.png"), img)
            cv2.imwrite(str(tmp_path / "targets" / f"t{i}_{kind}_disaster_target.png"), lbl)
    for mode, ch in (("pre", 3), ("post", 6)):
        loader = fetch_pytorch_loader(str(tmp_path), mode, True, dict(batch_size=2, shuffle=True, drop_last=True, num_workers=2, seed=1))
        batches = list(loader)
        assert len(batches) == 2 == len(loader)
        for b in batches:
            assert b["image"].shape == (2, ch, 512, 512) and b["image"].dtype == torch.bfloat16 and b["image"].is_cuda
            assert b["mask"].shape == (2, 512, 512) and b["mask"].dtype == torch.uint8
            assert bool(b["mask"].flatten(1).any(1).all()), "every crop must contain building pixels"
            assert float(b["image"].float().abs().max()) < 4.0  # normalised range
