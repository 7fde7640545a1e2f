"""CPU tests of the host side: CLI surface, tile sharding, loader datasets / augmentation, Noam schedule, flat gradient
all-reduce over gloo (world size 2), C-ABI symbol export.  No CUDA compute is called."""
import argparse
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import functional as OF  # noqa: E402


# ---------------------------------------------------------------------------------------------------------------
# C ABI
# ---------------------------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    from xview2_b200 import build, lib

    build.build()
    handle = lib.load()
    header = open(os.path.join(ROOT, "include", "xv2.h")).read()
    declared = set(re.findall(r"\b(xv2_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/xv2.h but not exported"
    assert set(lib.exported_symbols()) == declared


def test_ops_refuse_cpu_tensors():
    from xview2_b200 import lib, ops

    with pytest.raises(lib.Xv2Error):
        ops.conv2d(torch.zeros(1, 32, 8, 8), torch.zeros(32, 32, 3, 3), None, 1, 1)


# ---------------------------------------------------------------------------------------------------------------
# CLI (reference main.py:29-53 + plt.py:185-233)
# ---------------------------------------------------------------------------------------------------------------
def test_cli_flags_and_defaults_match_reference():
    sys.path.insert(0, ROOT)
    import main as cli

    a = cli.build_parser().parse_args(["--type", "pre"])
    expect = dict(exec_mode="train", data="/data", results="/results", gpus=1, num_workers=8, batch_size=16, val_batch_size=13,
                  precision=16, epochs=250, patience=100, ckpt=None, logname="logs", ckpt_pre=None, type="pre", seed=1,
                  optimizer="adamw", dmg_model="siamese", encoder="resnest200", loss_str="focal+dice", use_scheduler=False,
                  warmup=1, init_lr=1e-4, final_lr=1e-4, lr=3e-4, weight_decay=0, momentum=0.9, dilation=1, tta=False,
                  ppm=False, aspp=False, no_skip=False, deep_supervision=False, attention=False, autoaugment=False,
                  interpolate=False, dec_interp=False)
    for k, v in expect.items():
        assert getattr(a, k) == v, (k, getattr(a, k), v)
    assert len(vars(a)) == len(expect)
    assert cli.build_parser().parse_args(["--type", "post", "--precision", "bf16"]).precision == "bf16"
    with pytest.raises(SystemExit):
        cli.build_parser().parse_args(["--type", "both"])


def test_transplant_encoder_key_rules():
    import main as cli

    class Fake:
        def __init__(self, keys):
            self.sd = {k: torch.zeros(2) for k in keys}

        def state_dict(self):
            return self.sd

    pre = {"model.unet.enc_l1.0.0.weight": torch.ones(2), "model.unet.dec_l1.x": torch.ones(2)}
    m = Fake(["model.unet.enc_l1.0.0.weight", "model.unet.dec_l1.x"])
    assert cli.transplant_encoder(m, pre, "siamese") == 1
    assert m.sd["model.unet.enc_l1.0.0.weight"].sum() == 2 and m.sd["model.unet.dec_l1.x"].sum() == 0
    m = Fake(["model.enc_l1.0.0.weight"])
    assert cli.transplant_encoder(m, pre, "siameseEnc") == 1
    m = Fake(["model.unet_pre.enc_l1.0.0.weight", "model.unet_post.enc_l1.0.0.weight"])
    assert cli.transplant_encoder(m, pre, "parallel") == 2


# ---------------------------------------------------------------------------------------------------------------
# sharding = DistributedSampler semantics
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,world,shuffle", [(10, 2, False), (11, 4, True), (7, 8, False), (16, 1, True)])
def test_shard_indices_match_distributed_sampler(n, world, shuffle):
    from torch.utils.data.distributed import DistributedSampler

    from xview2_b200.data_loading.pytorch_loader import shard_indices

    seen = []
    for rank in range(world):
        ours = shard_indices(n, rank, world, shuffle, seed=3, epoch=2, drop_last=False, batch_size=1)
        if world > 1:
            ref = DistributedSampler(list(range(n)), num_replicas=world, rank=rank, shuffle=shuffle, seed=3)
            ref.set_epoch(2)
            assert ours == list(iter(ref))
        seen += ours
    assert set(seen) == set(range(n))  # every tile is visited; padding repeats only wrap around


# ---------------------------------------------------------------------------------------------------------------
# datasets + augmentation on synthetic PNG tiles
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture()
def tile_dir(tmp_path):
    import cv2

    rng = np.random.default_rng(0)
    for split in ("train", "test", "holdout"):
        os.makedirs(tmp_path / split / "images")
        os.makedirs(tmp_path / split / "targets")
        for i in range(3):
            for kind in ("pre", "post"):
                img = rng.integers(0, 256, (1024, 1024, 3), dtype=np.uint8)
                lbl = np.zeros((1024, 1024), np.uint8)
                lbl[300:420, 500:640] = 1 + (i % 4 if kind == "post" else 0)
                cv2.imwrite(str(tmp_path / split / "images" / f"tile{i}_{kind}_disaster.png"), img)
                cv2.imwrite(str(tmp_path / split / "targets" / f"tile{i}_{kind}_disaster_target.png"), lbl)
    return tmp_path


def test_datasets_return_decoded_uint8_tiles(tile_dir, monkeypatch):
    import cv2

    from xview2_b200.data_loading.pytorch_loader import TestDataset, TrainPostDataset, TrainPreDataset
    monkeypatch.setenv("XV2_HOST_AUG", "1")  # the host restatement of the augmentations (default: on the device, below)

    ds = TestDataset(str(tile_dir / "holdout"), "post", False)
    assert len(ds) == 3
    item = ds[1]
    ref = cv2.imread(str(tile_dir / "holdout" / "images" / "tile1_pre_disaster.png"))  # BGR, as the reference reads it
    assert item["tiles"].dtype == np.uint8 and item["tiles"].shape == (1024, 1024, 3) and np.array_equal(item["tiles"], ref)
    assert item["tiles_post"].shape == (1024, 1024, 3) and item["mask"].shape == (1024, 1024) and item["mask"].max() == 2
    pre = TrainPreDataset(str(tile_dir / "train"), "pre", False)
    a, b = pre[0], pre[0]
    assert a["tiles"].shape == (512, 512, 3) and a["mask"].shape == (512, 512)
    assert a["mask"].any(), "CropNonEmptyMaskIfExists must keep building pixels in the crop"
    assert np.array_equal(a["tiles"], b["tiles"]), "augmentation is a pure function of (seed, index, epoch)"
    post = TrainPostDataset(str(tile_dir / "train"), "post", False)
    item = post[2]
    assert item["tiles"].shape == item["tiles_post"].shape == (512, 512, 3) and item["mask"].any()
    with pytest.raises(NotImplementedError):
        TrainPreDataset(str(tile_dir / "train"), "pre", True)


def test_train_datasets_hand_decisions_to_the_device_augmentation(tile_dir):
    """Default training path: full decoded tiles + 19 host-drawn floats per sample; the pixels are augmented on the GPU."""
    from xview2_b200.data_loading.pytorch_loader import GpuTrainAugment, TrainPostDataset, TrainPreDataset

    pre = TrainPreDataset(str(tile_dir / "train"), "pre", False)
    assert pre.out_size == 1024
    a, b = pre[1], pre[1]
    assert a["tiles"].shape == (1024, 1024, 3) and a["mask"].shape == (1024, 1024)
    assert a["aug"].dtype == np.float32 and a["aug"].shape == (GpuTrainAugment.N_FLOATS,) and np.array_equal(a["aug"], b["aug"])
    post = TrainPostDataset(str(tile_dir / "train"), "post", False)
    assert post[0]["tiles_post"].shape == (1024, 1024, 3)
    import random
    draws = np.stack([GpuTrainAugment().draw(random.Random(s), 1024, 1024, 2) for s in range(4000)])
    zoom = draws[:, 15] == 1
    assert 0.17 < zoom.mean() < 0.23 and np.all(draws[zoom, 2] <= 1331) and np.all(draws[zoom, 2] >= 1024)
    assert np.allclose(draws[~zoom, :4], [1, 1, 1024, 1024])
    assert 0.30 < draws[:, 6].mean() < 0.36 and 0.30 < draws[:, 7].mean() < 0.36
    assert 0.08 < (draws[:, 8] > 0).mean() < 0.12 and 0.08 < (draws[:, 9] > 0).mean() < 0.12
    on = draws[:, 8] > 0
    assert np.all(draws[on, 8] ** 2 >= 10 - 1e-3) and np.all(draws[on, 8] ** 2 <= 50 + 1e-3)
    bc = (draws[:, 10] != 1) | (draws[:, 11] != 0)
    assert 0.17 < bc.mean() < 0.23 and np.all(np.abs(draws[:, 10] - 1) <= 0.2 + 1e-6) and np.all(np.abs(draws[:, 11]) <= 0.2 + 1e-6)
    # per-image independence (intensity_aug is called once per image, pytorch_loader.py:45-51)
    assert ((draws[:, 8] > 0) != (draws[:, 9] > 0)).mean() > 0.1


def test_augmentation_restatement_matches_cv2_resampling():
    """The float32 per-output-pixel restatement the device kernel is tested against reproduces cv2.resize (INTER_CUBIC for the
    image within one grey level on a vanishing fraction of pixels, INTER_NEAREST for the mask exactly)."""
    import cv2

    rng = np.random.default_rng(0)
    img = cv2.resize(rng.integers(0, 256, (64, 64, 3)).astype(np.uint8), (256, 256), interpolation=cv2.INTER_LINEAR)
    mask = (rng.random((256, 256)) > 0.97).astype(np.uint8)
    for scale in (1.07, 1.17, 1.3):
        ws, hs = int(256 * scale), int(256 * scale)
        P = np.zeros(16, np.float32)
        P[0], P[1], P[2], P[3], P[10], P[12], P[15] = 256 / ws, 256 / hs, ws, hs, 1, 1, 1
        ref = cv2.resize(img, (ws, hs), interpolation=cv2.INTER_CUBIC)
        refm = cv2.resize(mask, (ws, hs), interpolation=cv2.INTER_NEAREST)
        u8, _, m = OF.augment_restatement(img, None, mask, P, (10, 20), crop=128)
        d = np.abs(u8.astype(int) - ref[20:148, 10:138].astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 1e-3
        assert np.array_equal(m, refm[20:148, 10:138])
    # crop origin: the chosen crop always contains the selected non-zero pixel
    P[15] = 0
    P[:4] = [1, 1, 256, 256]
    yy, xx = np.nonzero(mask)
    for u in ((0.0, 0.0, 0.0), (0.5, 0.3, 0.9), (0.999, 0.999, 0.999)):
        px, py = OF.crop_origin_restatement(mask, P, u, crop=128)
        k = min(int(np.floor(np.float32(u[0]) * np.float32(len(yy)))), len(yy) - 1)
        assert 0 <= px <= 128 and 0 <= py <= 128 and px <= xx[k] < px + 128 and py <= yy[k] < py + 128


def test_normalize_constants_match_albumentations():
    """A.Normalize() (pytorch_loader.py:63): (x - 255*mean) * (1 / (255*std)) with the ImageNet constants, applied to BGR
    channel POSITIONS exactly as the reference does (cv2 loads BGR)."""
    from oracle import functional as OF

    x = np.arange(0, 256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    out = np.asarray(OF.normalize_tile(x))
    mean = np.array([0.485, 0.456, 0.406], np.float32) * 255
    rstd = 1 / (np.array([0.229, 0.224, 0.225], np.float32) * 255)
    ref = (x.astype(np.float32) - mean) * rstd
    assert np.allclose(out, np.transpose(ref, (2, 0, 1)), atol=1e-6)


def test_brightness_contrast_and_noise_are_uint8_safe():
    import random

    from xview2_b200.data_loading.pytorch_loader import TrainAugment

    img = np.full((64, 64, 6), 250, np.uint8)
    rng = random.Random(1)
    for _ in range(50):
        out = TrainAugment._brightness_contrast(rng, img)
        assert out.dtype == np.uint8 and out.shape == img.shape
    out = TrainAugment._noise(random.Random(5), np.random.default_rng(5), img)
    assert out.dtype == np.uint8


# ---------------------------------------------------------------------------------------------------------------
# Noam schedule (utils/scheduler.py:45-59)
# ---------------------------------------------------------------------------------------------------------------
def test_noam_lr_warmup_then_decay():
    from xview2_b200.utils.scheduler import NoamLR

    class Opt:
        param_groups = [{"lr": 0.0}]

    sch = NoamLR(Opt(), warmup_epochs=1, total_epochs=3, steps_per_epoch=10, init_lr=1e-4, max_lr=1e-3, final_lr=1e-5)
    # torch's _LRScheduler.__init__ (the reference's base class) steps once at construction: the first batch trains at step 1
    assert sch.current_step == 1 and abs(Opt.param_groups[0]["lr"] - (1e-4 + 9e-5)) < 1e-12
    lrs = [Opt.param_groups[0]["lr"]]
    for _ in range(30):
        sch.step()
        lrs.append(Opt.param_groups[0]["lr"])
    assert abs(lrs[9] - 1e-3) < 1e-9 and all(b > a for a, b in zip(lrs[:9], lrs[1:10]))
    assert all(b < a for a, b in zip(lrs[10:29], lrs[11:30])) and abs(lrs[-1] - 1e-5) < 1e-8
    for t, lr in enumerate(lrs, start=1):
        assert abs(lr - OF.noam_lr(t, 10, 30, 1e-4, 1e-3, 1e-5)) < 1e-12


@pytest.mark.skipif(not os.path.exists("/root/reference/utils/scheduler.py"), reason="reference tree not present")
def test_noam_lr_equals_reference_class_step_for_step():
    import importlib.util

    from xview2_b200.utils.scheduler import NoamLR

    spec = importlib.util.spec_from_file_location("ref_scheduler", "/root/reference/utils/scheduler.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    kw = dict(warmup_epochs=2, total_epochs=5, steps_per_epoch=7, init_lr=1e-4, max_lr=3e-4, final_lr=2e-5)
    p = torch.nn.Parameter(torch.zeros(1))
    ref_opt = torch.optim.SGD([p], lr=1.0)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref_sch = ref.NoamLR(optimizer=ref_opt, **kw)

        class Opt:
            param_groups = [{"lr": 0.0}]

        ours = NoamLR(Opt(), **kw)
        for _ in range(40):
            assert abs(Opt.param_groups[0]["lr"] - ref_opt.param_groups[0]["lr"]) < 1e-12
            ref_sch.step()
            ours.step()


def test_noam_steps_per_epoch_is_per_rank_once(monkeypatch):
    """TileLoader.__len__ is already per-rank; configure_optimizers must not divide by --gpus again."""
    from xview2_b200.model import plt as xplt

    class FakeFlat:
        params = []

    class FakeModel:
        args = argparse.Namespace(optimizer="adamw", weight_decay=0.0, use_scheduler=True, warmup=1, epochs=4, gpus=4,
                                  init_lr=1e-4, final_lr=1e-5, lr=3e-4, momentum=0.9)
        lr = 3e-4
        flat = FakeFlat()

        def train_dataloader(self):
            return range(25)  # 25 batches on THIS rank

    made = {}

    class FakeOpt:
        def __init__(self, flat, **kw):
            self.param_groups = [{"lr": kw["lr"]}]

    monkeypatch.setattr(xplt, "FusedAdamW", FakeOpt)
    conf = xplt.Model.configure_optimizers(FakeModel())
    sch = conf["lr_scheduler"]["scheduler"]
    assert sch.steps_per_epoch == 25 and sch.warmup_steps == 25 and sch.total_steps == 100


def test_reference_checkpoint_with_metric_states_loads_strictly():
    """pytorch_lightning 1.0 Metric states are persistent: reference checkpoints carry f1_score.tp/fp/fn."""
    from xview2_b200.model.plt import Model

    ns = argparse.Namespace(ppm=False, aspp=False, dilation=1, no_skip=False, interpolate=False, attention=False,
                            dec_interp=False, deep_supervision=False, loss_str="focal+dice", encoder="resnet50",
                            dmg_model="siamese", type="post", tta=False, precision="bf16", lr=3e-4, results=None, logname="t")
    m = Model(ns)
    ck = m.checkpoint()
    assert {"f1_score.tp", "f1_score.fp", "f1_score.fn"} <= set(ck["state_dict"])
    sd = dict(ck["state_dict"])
    sd["f1_score.tp"] = torch.tensor([1.0, 2.0, 3.0, 4.0])
    sd["f1_score.fp"] = torch.tensor([5.0, 6.0, 7.0, 8.0])
    sd["f1_score.fn"] = torch.zeros(4)
    m2 = Model(ns).load_reference_state_dict(sd)
    assert m2.f1_score.tp.tolist() == [1.0, 2.0, 3.0, 4.0] and m2.f1_score.fp.tolist() == [5.0, 6.0, 7.0, 8.0]
    with pytest.raises(RuntimeError):  # still strict for everything else
        m2.load_reference_state_dict({**sd, "model.bogus": torch.zeros(1)})


def test_checkpoint_callback_saves_scheduler_and_global_step(tmp_path):
    from xview2_b200.trainer import ModelCheckpoint
    from xview2_b200.utils.scheduler import NoamLR

    class Opt:
        param_groups = [{"lr": 0.0}]

        def state_dict(self):
            return {"step": 3}

    class FakeModel:
        logged = {"f1_score": torch.tensor(0.5)}
        current_epoch = 2

        def checkpoint(self):
            return {"state_dict": {}, "hyper_parameters": {}, "epoch": 2}

    class FakeTrainer:
        global_rank = 0
        default_root_dir = str(tmp_path)
        global_step = 77
        scheduler = NoamLR(Opt(), 1, 3, 10, 1e-4, 1e-3, 1e-5)

    FakeTrainer.scheduler.step(33)
    cb = ModelCheckpoint(dirpath=str(tmp_path))
    cb.on_epoch_end(FakeTrainer(), FakeModel(), Opt())
    ck = torch.load(os.path.join(str(tmp_path), "last.ckpt"), weights_only=False)
    assert ck["global_step"] == 77 and ck["lr_schedulers"][0]["current_step"] == 33
    fresh = NoamLR(Opt(), 1, 3, 10, 1e-4, 1e-3, 1e-5)
    fresh.load_state_dict(ck["lr_schedulers"][0])
    assert fresh.current_step == 33 and abs(Opt.param_groups[0]["lr"] - FakeTrainer.scheduler.lr[0]) < 1e-15


def test_intensity_augmentations_draw_per_image():
    """intensity_aug (pytorch_loader.py:45-51) applies GaussNoise / RandomBrightnessContrast once per 3-channel image."""
    import random

    from xview2_b200.data_loading.pytorch_loader import TrainAugment

    img = np.full((32, 32, 6), 100, np.uint8)
    differ = 0
    for seed in range(200):
        out = TrainAugment._brightness_contrast(random.Random(seed), img)
        a, b = out[:, :, :3], out[:, :, 3:]
        differ += int(a[0, 0, 0] != b[0, 0, 0])
    assert differ > 20  # with one shared draw the two halves would always be equal


# ---------------------------------------------------------------------------------------------------------------
# data-parallel plumbing over gloo, world size 2: ONE all-reduce of the flat gradient buffer + rank-0 broadcast
# ---------------------------------------------------------------------------------------------------------------
def _dp_worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from xview2_b200.data_loading.pytorch_loader import shard_indices
    from xview2_b200.optim import FlatParams

    torch.manual_seed(100 + rank)  # different initial weights per rank: the broadcast must make them equal
    net = torch.nn.Sequential(torch.nn.Conv2d(4, 8, 3, bias=False), torch.nn.BatchNorm2d(8), torch.nn.Conv2d(8, 2, 1))
    flat = FlatParams(net)
    flat.broadcast_params(0)
    ref = flat.data.clone()
    dist.broadcast(ref, 0)
    same_params = bool(torch.equal(ref, flat.data))
    # tiles sharded over ranks: the per-rank "gradient" is the sum of its tile indices
    tiles = shard_indices(10, rank, world, False, 1, 0, False, 1)
    for p in net.parameters():
        p.grad.fill_(float(sum(tiles)))
    n = flat.all_reduce_grads()
    views_alias_flat = all(p.grad.data_ptr() >= flat.grad.data_ptr() for p in net.parameters())
    out[rank] = (same_params, n, float(flat.grad[0]), views_alias_flat)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo_world2():
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    out = mgr.dict()
    port = 29600 + os.getpid() % 200
    mp.spawn(_dp_worker, args=(2, port, out), nprocs=2, join=True)
    assert len(out) == 2
    for rank in range(2):
        same, n, g0, alias = out[rank]
        assert same and n == 2 and alias
        assert g0 == float(sum(range(10)))  # SUM over ranks of disjoint tile shards = sum over all tiles


def _bucket_worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from xview2_b200.optim import FlatParams

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.enc_l1 = torch.nn.Conv2d(3, 4, 3, padding=1)
            self.enc_l2 = torch.nn.Conv2d(4, 4, 3, padding=1)
            self.enc_l3 = torch.nn.Conv2d(4, 4, 3, padding=1)
            self.dec_l1 = torch.nn.Conv2d(8, 4, 3, padding=1)
            self.head = torch.nn.Conv2d(4, 2, 1)

        def forward(self, x):
            e1 = self.enc_l1(x)
            e2 = self.enc_l2(e1)
            e3 = self.enc_l3(e2) + self.enc_l3(e1)  # a stage visited twice (Siamese-style sharing)
            return self.head(self.dec_l1(torch.cat((e3, e2), 1)))

    torch.manual_seed(7)
    net = Net()
    flat = FlatParams(net)
    nb = flat.enable_bucketed_allreduce(net, min_bytes=0)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(2, 3, 8, 8, generator=g)
    early = []
    orig = flat._stage_done

    def spy(b):
        orig(b)
        early.append((b["lo"], b["done"]))
    flat._stage_done = spy
    for b in flat._buckets:  # re-register with the spy
        pass
    flat.zero_grad()
    net(x).sum().backward()
    launched_in_backward = sum(1 for b in flat._buckets if b["done"])
    local = None
    n = flat.all_reduce_grads()
    got = flat.grad.clone()
    # reference: un-bucketed all-reduce of the same local gradients
    flat.disable_bucketed_allreduce()
    flat.zero_grad()
    net(x).sum().backward()
    flat.all_reduce_grads()
    out[rank] = (nb, launched_in_backward, n, bool(torch.allclose(got, flat.grad, rtol=1e-6, atol=1e-7)), float(got.abs().sum()))
    dist.destroy_process_group()


def test_bucketed_allreduce_overlaps_backward_and_matches_flat_gloo_world2():
    """DDP-style buckets on the flat gradient buffer: per-stage collectives are launched from backward hooks (a twice-visited
    stage only after its LAST visit), the remainder by all_reduce_grads(); the result equals one flat all-reduce."""
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    out = mgr.dict()
    port = 29800 + os.getpid() % 150
    mp.spawn(_bucket_worker, args=(2, port, out), nprocs=2, join=True)
    assert len(out) == 2
    for rank in range(2):
        nb, launched, n, same, mag = out[rank]
        assert nb >= 3 and launched >= 2 and n == 2 and same and mag > 0, out[rank]


# ---------------------------------------------------------------------------------------------------------------
# host steps of utils/post_process.py (connected-component vote, dilation) against the reference's own formulation
# ---------------------------------------------------------------------------------------------------------------
def test_majority_vote_and_dilation_match_reference_formulation():
    from scipy.ndimage import label

    from xview2_b200.utils.post_process import dilate, majority_vote

    rng = np.random.default_rng(4)
    cells = rng.integers(0, 5, (24, 24))
    cells[rng.random((24, 24)) < 0.45] = 0
    post = np.kron(cells, np.ones((8, 8), dtype=np.int64)).astype(np.uint8)
    noise = rng.random(post.shape) < 0.2
    post[noise & (post > 0)] = rng.integers(1, 5, int((noise & (post > 0)).sum()))
    ref = post.copy()
    components, n = label(ref > 0)  # the reference's loop, post_process.py:39-43
    for b in range(1, n + 1):
        labels, counts = np.unique(ref[components == b], return_counts=True)
        ref[components == b] = labels[np.argmax(counts)]
    assert n > 5 and np.array_equal(majority_vote(post), ref)
    assert np.array_equal(majority_vote(np.zeros((8, 8), np.uint8)), np.zeros((8, 8), np.uint8))
    # square grey dilation == running maximum over the k x k window clipped at the border (skimage dilation(img, square(k)))
    for k in (3, 5):
        out = dilate(post, k)
        pad = k // 2
        padded = np.pad(post, pad, constant_values=0)
        want = np.max([padded[i:i + post.shape[0], j:j + post.shape[1]] for i in range(k) for j in range(k)], axis=0)
        assert np.array_equal(out, want)


# ---------------------------------------------------------------------------------------------------------------
# drop-in contract: state_dict keys and shapes of EVERY model variant == the reference's own modules
# (tests/golden/state_keys.json, written by tools/make_state_keys.py from /root/reference/model/unet.py)
# ---------------------------------------------------------------------------------------------------------------
def _state_key_cases():
    import json

    with open(os.path.join(ROOT, "tests", "golden", "state_keys.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", sorted(_state_key_cases()))
def test_state_dict_keys_match_reference(name):
    from xview2_b200.model.unet import UNetLoc, get_dmg_unet

    case = _state_key_cases()[name]
    ns = argparse.Namespace(**case["args"])
    model = UNetLoc(ns) if ns.type == "pre" else get_dmg_unet(ns)
    ours = {k: list(v.shape) for k, v in model.state_dict().items()}
    ref = case["keys"]
    missing = sorted(set(ref) - set(ours))
    extra = sorted(set(ours) - set(ref))
    assert not missing and not extra, (missing[:5], extra[:5])
    wrong = [(k, ours[k], ref[k]) for k in ref if ours[k] != ref[k]]
    assert not wrong, wrong[:5]


# ---------------------------------------------------------------------------------------------------------------
# gradient plumbing of the captured step (host logic only)
# ---------------------------------------------------------------------------------------------------------------
def test_fork_parks_second_gradient_part_and_unclaimed_parts_fail_loudly():
    """ops.fork hands one gradient part on and parks the other for the producer's BatchNorm backward; if nobody collects it
    the step-boundary check raises instead of training on a silently incomplete gradient."""
    from xview2_b200 import ops

    x = torch.randn(2, 8, 4, 4, requires_grad=True)
    y = x * 1.0
    assert ops.fork(y)[0] is y and ops.fork(y)[1] is y      # not a train-mode BatchNorm output: plain aliases, autograd accumulates
    y._xv2_forkable = True
    a, b = ops.fork(y)
    assert a is not y and a.data_ptr() == y.data_ptr() and b.data_ptr() == y.data_ptr()
    (a.sum() + 2.0 * b.sum()).backward()                     # the consumer here is a plain multiply: nobody pops the parked part
    assert torch.equal(x.grad, torch.ones_like(x))           # only the first part travelled on
    with pytest.raises(RuntimeError, match="never consumed"):
        ops.check_pending_addends()
    ops.check_pending_addends()                              # the table is cleared by the failure
    with torch.no_grad():
        assert ops.fork(y)[0] is y                           # no tape: no fork


def test_deferred_num_batches_tracked_counts_every_visit():
    """ops.defer_nbt: the per-BatchNorm `num_batches_tracked += 1` launches of a forward are applied as one multi-tensor add per
    multiplicity (a module visited twice -- Siamese encoders -- is bumped by two)."""
    from xview2_b200 import ops

    bns = [torch.nn.BatchNorm2d(4) for _ in range(3)]
    with ops.defer_nbt():
        for bn in (bns[0], bns[1], bns[0], bns[2], bns[0]):
            ops._bump_nbt(bn)
        assert all(int(bn.num_batches_tracked) == 0 for bn in bns)   # nothing applied yet
    assert [int(bn.num_batches_tracked) for bn in bns] == [3, 1, 1]
    ops._bump_nbt(bns[1])                                              # outside the context: immediate
    assert int(bns[1].num_batches_tracked) == 2
    with pytest.raises(ValueError):
        with ops.defer_nbt():
            ops._bump_nbt(bns[2])
            raise ValueError("forward failed")
    assert int(bns[2].num_batches_tracked) == 1                        # a failed forward applies nothing


def test_every_abi_declaration_sits_under_a_reference_citation():
    """include/xv2.h: each entry point is declared under a comment (its own or its section's) that cites the reference call site
    it replaces as file.py:line -- the drop-in boundary is documented where it is declared."""
    lines = open(os.path.join(ROOT, "include", "xv2.h")).read().splitlines()
    cite = re.compile(r"[a-z_0-9]+\.py:\d+")
    housekeeping = {"xv2_last_error", "xv2_version", "xv2_init"}
    seen = 0
    for i, line in enumerate(lines):
        m = re.match(r"(?:int|const char\*) (xv2_[a-z0-9_]+)\s*\(", line)
        if not m or m.group(1) in housekeeping:
            continue
        seen += 1
        assert cite.search("\n".join(lines[max(0, i - 60):i])), f"{m.group(1)}: no reference file:line above its declaration"
    assert seen >= 70
