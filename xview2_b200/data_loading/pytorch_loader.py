"""Tile -> GPU loader with the reference's names (/root/reference/data_loading/pytorch_loader.py): ``fetch_pytorch_loader``,
``load_data``, ``load_pair``, ``TrainPreDataset``, ``TrainPostDataset``, ``TestDataset``, ``seed_worker``.

What differs from the reference (by design):
  * datasets return the DECODED uint8 tiles (cv2 BGR HWC, exactly what ``cv2.imread`` gives, pytorch_loader.py:39-42) and the
    uint8 mask; ``A.Normalize`` + HWC->CHW (pytorch_loader.py:63,91) run on the GPU in one kernel (xv2_normalize_tiles), so
    3 B/pixel cross PCIe instead of 12;
  * ``fetch_pytorch_loader`` returns a ``TileLoader``: decode THREADS (cv2 releases the GIL) write straight into the slots of a
    pinned-host ring (xview2_b200.data_loading.ring.TileRing); the H2D copy of batch i+1 runs on a side stream under the
    compute of batch i.  Batches are dicts of DEVICE tensors {"tiles", ["tiles_post"], "mask"} that ``Model._image`` accepts;
  * tiles are sharded over data-parallel ranks with DistributedSampler semantics (index i -> rank i mod N, padded by wrap-around).

The train-time augmentations (albumentations 0.5.1, not installed here: RandomScale(p=.2, 1.0-1.3x, cubic) ->
CropNonEmptyMaskIfExists(512) -> H/V flip (p=.33) -> GaussNoise(p=.1) -> RandomBrightnessContrast(p=.2) -> Normalize) run ON THE
DEVICE by default: the training datasets hand over the full decoded tile plus a dozen host-drawn decisions per sample
(``GpuTrainAugment``), and one gather kernel per batch (xv2_augment_tiles; crop origin from the mask content by xv2_crop_origin)
writes the normalised 512^2 crops.  ``XV2_HOST_AUG=1`` selects the host restatement (``TrainAugment``, uint8 numpy / cv2)
instead.  ``--autoaugment`` (PIL ImageNet policy) is outside the accelerated path.
"""
import os
import random
from concurrent.futures import ThreadPoolExecutor
from glob import glob

import numpy as np
import torch

from .ring import TileRing

INDEX_CSV_CANDIDATES = ("/workspace/xview2/utils/index.csv",)  # the reference's hard-coded path (pytorch_loader.py:64,101)


def seed_worker(worker_id):
    worker_seed = torch.initial_seed() % 2 ** 32
    np.random.seed(worker_seed)
    random.seed(worker_seed)


def load_data(path, dtype):
    imgs = sorted(glob(os.path.join(path, "images", f"*{dtype}*")))
    lbls = sorted(glob(os.path.join(path, "targets", f"*{dtype}*")))
    assert len(imgs) == len(lbls) and len(imgs) > 0
    return imgs, lbls


def load_pair(img, lbl):
    import cv2

    img = cv2.imread(img)
    lbl = cv2.imread(lbl, cv2.IMREAD_UNCHANGED)
    return img, lbl


def _read_index(path):
    """utils/index.csv of the reference: columns idx, 1, 2, 3, 4.  Looked up at $XVIEW2_INDEX_CSV, the reference's hard-coded
    location, then next to the data directory.  Returns None when absent (every tile is then used)."""
    cands = [os.environ.get("XVIEW2_INDEX_CSV")] + list(INDEX_CSV_CANDIDATES) + [
        os.path.join(os.path.dirname(os.path.abspath(path)), "index.csv"), os.path.join(path, "index.csv")]
    for c in cands:
        if c and os.path.exists(c):
            import pandas as pd

            return pd.read_csv(c)
    return None


# ---------------------------------------------------------------------------------------------------------------
# augmentations (uint8 host arrays; albumentations 0.5.1 semantics)
# ---------------------------------------------------------------------------------------------------------------
class TrainAugment:
    def __init__(self, crop=512):
        self.crop = crop

    @staticmethod
    def _zoom(rng, img, lbl):
        import cv2

        if rng.random() >= 0.2:
            return img, lbl
        scale = rng.uniform(1.0, 1.3)
        h, w = lbl.shape[:2]
        size = (int(w * scale), int(h * scale))
        chans = [cv2.resize(np.ascontiguousarray(img[:, :, i:i + 3]), size, interpolation=cv2.INTER_CUBIC)
                 for i in range(0, img.shape[2], 3)]
        return np.concatenate(chans, 2), cv2.resize(lbl, size, interpolation=cv2.INTER_NEAREST)

    def _crop(self, rng, img, lbl):
        ch = cw = self.crop
        h, w = lbl.shape[:2]
        if lbl.any():
            ys, xs = np.nonzero(lbl if lbl.ndim == 2 else lbl.sum(-1))
            j = rng.randrange(len(ys))
            y_min = int(np.clip(ys[j] - rng.randint(0, ch - 1), 0, h - ch))
            x_min = int(np.clip(xs[j] - rng.randint(0, cw - 1), 0, w - cw))
        else:
            y_min, x_min = rng.randint(0, h - ch), rng.randint(0, w - cw)
        return img[y_min:y_min + ch, x_min:x_min + cw], lbl[y_min:y_min + ch, x_min:x_min + cw]

    @staticmethod
    def _noise(rng, nprng, img):
        """A.GaussNoise(p=0.1): intensity_aug (pytorch_loader.py:45-51) calls the transform once PER IMAGE, so the pre and the
        post tile each get their own probability draw, variance and noise field."""
        out = []
        for i in range(0, img.shape[2], 3):
            part = img[:, :, i:i + 3]
            if rng.random() < 0.1:
                sigma = rng.uniform(10.0, 50.0) ** 0.5
                g = nprng.normal(0.0, sigma, part.shape)
                part = np.clip(part.astype(np.float32) + g, 0, 255).astype(np.uint8)
            out.append(part)
        return out[0] if len(out) == 1 else np.concatenate(out, 2)

    @staticmethod
    def _brightness_contrast(rng, img):
        """A.RandomBrightnessContrast(p=0.2), likewise one independent draw (p, alpha, beta) per 3-channel image."""
        out = []
        for i in range(0, img.shape[2], 3):
            part = img[:, :, i:i + 3]
            if rng.random() < 0.2:
                alpha, beta = 1.0 + rng.uniform(-0.2, 0.2), rng.uniform(-0.2, 0.2)
                lut = np.arange(0, 256, dtype=np.float32) * alpha
                if beta != 0:
                    lut += beta * 255.0
                part = np.clip(lut, 0, 255).astype(np.uint8)[part]
            out.append(part)
        return out[0] if len(out) == 1 else np.concatenate(out, 2)

    def __call__(self, rng, nprng, img, lbl):
        img, lbl = self._zoom(rng, img, lbl)
        img, lbl = self._crop(rng, img, lbl)
        if rng.random() < 0.33:
            img, lbl = img[:, ::-1], lbl[:, ::-1]
        if rng.random() < 0.33:
            img, lbl = img[::-1], lbl[::-1]
        img = self._noise(rng, nprng, img)
        img = self._brightness_contrast(rng, img)
        return np.ascontiguousarray(img), np.ascontiguousarray(lbl)


class GpuTrainAugment:
    """Host half of the device-side augmentation: draws the per-sample DECISIONS (19 floats: the 16-float parameter block of
    xv2_augment_tiles + 3 uniforms for the crop origin); every pixel is touched on the GPU only."""

    N_FLOATS = 19

    def __init__(self, crop=512):
        self.crop = crop

    def draw(self, rng, h, w, n_images):
        p = np.zeros(self.N_FLOATS, np.float32)
        ws, hs = w, h
        if rng.random() < 0.2:  # A.RandomScale(p=0.2, scale_limit=(0, 0.3)): cv2.resize to int(size * scale)
            scale = rng.uniform(1.0, 1.3)
            ws, hs = int(w * scale), int(h * scale)
            p[15] = 1.0
        p[0], p[1], p[2], p[3] = w / ws, h / hs, ws, hs
        p[6] = float(rng.random() < 0.33)   # A.HorizontalFlip(p=0.33)
        p[7] = float(rng.random() < 0.33)   # A.VerticalFlip(p=0.33)
        for im in range(2):
            on = im < n_images
            p[8 + im] = rng.uniform(10.0, 50.0) ** 0.5 if (on and rng.random() < 0.1) else 0.0   # A.GaussNoise(p=0.1), per image
            if on and rng.random() < 0.2:                                                        # A.RandomBrightnessContrast(p=0.2)
                p[10 + 2 * im], p[11 + 2 * im] = 1.0 + rng.uniform(-0.2, 0.2), rng.uniform(-0.2, 0.2)
            else:
                p[10 + 2 * im], p[11 + 2 * im] = 1.0, 0.0
        p[14] = float(rng.randrange(1 << 24))
        p[16:19] = [rng.random(), rng.random(), rng.random()]
        return p


def gpu_augment_enabled():
    return os.environ.get("XV2_HOST_AUG", "0") != "1"


# ---------------------------------------------------------------------------------------------------------------
# datasets: __getitem__ -> {"tiles": u8 HxWx3 (BGR), ["tiles_post": u8 HxWx3], "mask": u8 HxW}
# ---------------------------------------------------------------------------------------------------------------
class _Dataset:
    out_size = 1024

    def __len__(self):
        return len(self.idx)

    def _rngs(self, idx):
        seed = (torch.initial_seed() + 7919 * idx + 104729 * getattr(self, "epoch", 0)) % 2 ** 32
        return random.Random(seed), np.random.default_rng(seed)


class TrainPreDataset(_Dataset):
    out_size = 512

    def __init__(self, path, _, autoaugment):
        if autoaugment:
            raise NotImplementedError("--autoaugment (PIL ImageNet policy, autoaugment.py) is outside the accelerated path")
        self.imgs_pre, self.lbls_pre = load_data(path, "pre")
        frame = _read_index(path)
        self.idx = frame["idx"].tolist() if frame is not None else list(range(len(self.imgs_pre)))
        self.aug = TrainAugment(512)
        self.gpu_aug = GpuTrainAugment(512) if gpu_augment_enabled() else None
        if self.gpu_aug is not None:
            self.out_size = 1024  # the ring carries the full decoded tile; the 512^2 crop is cut on the device

    def __getitem__(self, idx):
        img, lbl = load_pair(self.imgs_pre[self.idx[idx]], self.lbls_pre[self.idx[idx]])
        if self.gpu_aug is not None:
            rng, _ = self._rngs(idx)
            return {"tiles": img, "mask": lbl, "aug": self.gpu_aug.draw(rng, lbl.shape[0], lbl.shape[1], 1)}
        img, lbl = self.aug(*self._rngs(idx), img, lbl)
        return {"tiles": img, "mask": lbl}


class TrainPostDataset(_Dataset):
    out_size = 512

    def __init__(self, path, _, autoaugment):
        if autoaugment:
            raise NotImplementedError("--autoaugment (PIL ImageNet policy, autoaugment.py) is outside the accelerated path")
        self.imgs_pre, self.lbls_pre = load_data(path, "pre")
        self.imgs_post, self.lbls_post = load_data(path, "post")
        assert len(self.imgs_pre) == len(self.imgs_post)
        assert len(self.imgs_post) == len(self.lbls_post)
        frame = _read_index(path)
        if frame is not None:  # tiles that contain at least one damage class (pytorch_loader.py:101-107)
            keep = set()
            for col in ("1", "2", "3", "4"):
                keep.update(frame[frame[col] == 1]["idx"].values.tolist())
            self.idx = sorted(keep)
        else:
            self.idx = list(range(len(self.imgs_pre)))
        self.aug = TrainAugment(512)
        self.gpu_aug = GpuTrainAugment(512) if gpu_augment_enabled() else None
        if self.gpu_aug is not None:
            self.out_size = 1024

    def __getitem__(self, idx):
        img_pre, _ = load_pair(self.imgs_pre[self.idx[idx]], self.lbls_pre[self.idx[idx]])
        img_post, lbl = load_pair(self.imgs_post[self.idx[idx]], self.lbls_post[self.idx[idx]])
        if self.gpu_aug is not None:
            rng, _ = self._rngs(idx)
            return {"tiles": img_pre, "tiles_post": img_post, "mask": lbl,
                    "aug": self.gpu_aug.draw(rng, lbl.shape[0], lbl.shape[1], 2)}
        img, lbl = self.aug(*self._rngs(idx), np.concatenate((img_pre, img_post), axis=2), lbl)
        return {"tiles": np.ascontiguousarray(img[:, :, :3]), "tiles_post": np.ascontiguousarray(img[:, :, 3:]), "mask": lbl}


class TestDataset(_Dataset):
    __test__ = False  # not a pytest class

    def __init__(self, path, mode, _):
        self.mode = mode
        self.imgs_pre, self.lbls_pre = load_data(path, "pre")
        self.imgs_post, self.lbls_post = load_data(path, "post")
        assert len(self.imgs_pre) == len(self.imgs_post)
        assert len(self.imgs_post) == len(self.lbls_post)
        self.idx = list(range(len(self.imgs_pre)))

    def __getitem__(self, idx):
        img, lbl = load_pair(self.imgs_pre[idx], self.lbls_pre[idx])
        if self.mode == "post":
            img_post, lbl = load_pair(self.imgs_post[idx], self.lbls_post[idx])
            return {"tiles": img, "tiles_post": img_post, "mask": lbl}
        return {"tiles": img, "mask": lbl}


# ---------------------------------------------------------------------------------------------------------------
# loader
# ---------------------------------------------------------------------------------------------------------------
def shard_indices(n, rank, world, shuffle, seed, epoch, drop_last, batch_size):
    """DistributedSampler semantics: seeded permutation (train), padded by wrap-around to a multiple of `world`, rank r takes
    positions r, r+world, ...; then whole batches only when drop_last."""
    order = list(range(n))
    if shuffle:
        g = torch.Generator().manual_seed(seed + epoch)
        order = torch.randperm(n, generator=g).tolist()
    if world > 1:
        total = -(-n // world) * world
        order = (order + order[:total - n])[:total] if n else order
        order = order[rank:total:world]
    if drop_last:
        order = order[:len(order) // batch_size * batch_size]
    return order


class TileLoader:
    """Iterable over device batches; see the module docstring.  ``len()`` = batches this rank yields per epoch."""

    def __init__(self, dataset, batch_size, shuffle=False, drop_last=False, num_workers=8, pin_memory=True, seed=1,
                 rank=None, world=None, device=None, slots=3):
        self.dataset, self.batch_size, self.shuffle, self.drop_last = dataset, batch_size, shuffle, drop_last
        self.num_workers = max(1, num_workers)
        self.seed, self.epoch = seed, 0
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else rank
        self.world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
        self.device, self.slots = device, slots
        self._ring = None

    def set_epoch(self, epoch):
        self.epoch = epoch
        self.dataset.epoch = epoch

    def _order(self):
        return shard_indices(len(self.dataset), self.rank, self.world, self.shuffle, self.seed, self.epoch, self.drop_last,
                             self.batch_size)

    def __len__(self):
        n = len(self._order())
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def _fill(self, slot, j, index):
        item = self.dataset[index]
        for k, v in item.items():
            slot[k][j].copy_(torch.from_numpy(v))

    def __iter__(self):
        order = self._order()
        batches = [order[i:i + self.batch_size] for i in range(0, len(order), self.batch_size)]
        if not batches:
            return
        size = self.dataset.out_size
        post = getattr(self.dataset, "mode", "pre") == "post" or isinstance(self.dataset, TrainPostDataset)
        gpu_aug = getattr(self.dataset, "gpu_aug", None)
        if self._ring is None or self._ring.batch != self.batch_size or self._ring.hw != (size, size):
            extra = {"aug": ((GpuTrainAugment.N_FLOATS,), torch.float32)} if gpu_aug is not None else None
            self._ring = TileRing(self.batch_size, size, size, post=post, slots=self.slots, device=self.device, extra=extra)
            self._ring.batch, self._ring.hw = self.batch_size, (size, size)
        ring = self._ring
        with ThreadPoolExecutor(self.num_workers) as pool:
            def decode(i):
                slot = ring.host(i)
                return [pool.submit(self._fill, slot, j, idx) for j, idx in enumerate(batches[i])]

            ahead = min(self.slots - 1, len(batches))
            pending = {}
            for i in range(ahead):
                ring.wait_free(i)
                pending[i] = decode(i)
            for i in range(len(batches)):
                for f in pending.pop(i):
                    f.result()
                ring.submit(i)
                nxt = i + ahead
                if nxt < len(batches):
                    ring.wait_free(nxt)  # host: the H2D copy that last read this slot has finished
                    pending[nxt] = decode(nxt)
                dev = ring.acquire(i)
                n = len(batches[i])
                if gpu_aug is not None:
                    # the whole augmentation chain + Normalize as one gather kernel on the uploaded uint8 tiles
                    from .. import ops
                    aug = dev["aug"][:n]
                    image, mask, _ = ops.augment_tiles(dev["tiles"][:n], dev["tiles_post"][:n] if "tiles_post" in dev else None,
                                                       dev["mask"][:n], aug[:, :16], aug[:, 16:19], crop=gpu_aug.crop)
                    ring.release(i)  # the crops are new tensors: the slot may be refilled as soon as the kernel has run
                    yield {"image": image, "mask": mask}
                    continue
                yield {k: v[:n] for k, v in dev.items()}
                ring.release(i)


def fetch_pytorch_loader(path, mode, training, loader_kwargs, autoaugment=False):
    """pytorch_loader.py:22-29.  ``loader_kwargs`` are the DataLoader kwargs the reference's DataModule builds
    (batch_size, pin_memory, num_workers, drop_last, shuffle)."""
    if not training:
        dataset = TestDataset
    elif mode == "pre":
        dataset = TrainPreDataset
    else:
        dataset = TrainPostDataset
    return TileLoader(dataset(path, mode, autoaugment), **loader_kwargs)
