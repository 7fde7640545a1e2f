"""End-to-end GPU test of the host side around the kernels: synthetic PNG tiles on disk -> DataModule / TileLoader (decode
threads, pinned ring, side-stream H2D, GPU normalise) -> Trainer.fit (training_step / backward / flat all-reduce / fused
AdamW / Noam) -> validation (TTA forward, F1) -> checkpoint -> Model.load_from_checkpoint -> Trainer.test -> .npy
probabilities as the reference writes them -> post-process label maps (main.py:96-122, plt.py:50-144, post_process.py:27-38)."""
import argparse
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _write_tiles(root, n_tiles=3):
    import cv2

    rng = np.random.default_rng(0)
    for split in ("train", "test", "holdout"):
        os.makedirs(os.path.join(root, split, "images"))
        os.makedirs(os.path.join(root, split, "targets"))
        for i in range(n_tiles):
            base = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
            for kind in ("pre", "post"):
                img = cv2.resize(base, (1024, 1024), interpolation=cv2.INTER_LINEAR) + (5 if kind == "post" else 0)
                lbl = np.zeros((1024, 1024), np.uint8)
                lbl[200 + 100 * i:420 + 100 * i, 300:700] = 1 + (i % 4 if kind == "post" else 0)
                cv2.imwrite(os.path.join(root, split, "images", f"tile{i}_{kind}_disaster.png"), img.astype(np.uint8))
                cv2.imwrite(os.path.join(root, split, "targets", f"tile{i}_{kind}_disaster_target.png"), lbl)


def _args(data, results, task="pre", **kw):
    import main as cli

    argv = ["--type", task, "--dmg_model", "siamese", "--encoder", "resnest50", "--data", data, "--results", results, "--batch_size", "2",
            "--val_batch_size", "2", "--num_workers", "4", "--epochs", "1", "--precision", "bf16", "--use_scheduler", "--tta"]
    a = cli.build_parser().parse_args(argv)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


@pytest.mark.parametrize("task", ["pre", "post"])
def test_fit_checkpoint_test_postprocess(tmp_path, monkeypatch, task):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from xview2_b200 import ops
    from xview2_b200.data_loading.data_module import DataModule
    from xview2_b200.model.plt import Model
    from xview2_b200.trainer import EarlyStopping, ModelCheckpoint, Trainer, seed_everything

    data, results = str(tmp_path / "data"), str(tmp_path / "results")
    _write_tiles(data)
    args = _args(data, results, task)
    seed_everything(args.seed)
    dm = DataModule(args)
    # loader alone: batches are device uint8 tiles the GPU normalises
    first = next(iter(dm.val_dataloader()))
    assert first["tiles"].is_cuda and first["tiles"].dtype == torch.uint8 and first["tiles"].shape == (2, 1024, 1024, 3)
    assert first["mask"].shape == (2, 1024, 1024) and int(first["mask"].max()) == (1 if task == "pre" else 2)
    assert ("tiles_post" in first) == (task == "post")
    img = Model._image(first)
    assert img.shape[1] == (3 if task == "pre" else 6)
    ref = (first["tiles"][0, 100, 200].float().cpu() - torch.tensor([0.485, 0.456, 0.406]) * 255) / (torch.tensor([0.229, 0.224, 0.225]) * 255)
    assert torch.allclose(img[0, :3, 100, 200].float().cpu(), ref, atol=2e-2)

    model = Model(args)
    ckpt_cb = ModelCheckpoint(monitor="f1_score", mode="max", save_last=True)
    trainer = Trainer(gpus=1, precision=args.precision, callbacks=[EarlyStopping(patience=3)], max_epochs=1, min_epochs=1,
                      default_root_dir=results, checkpoint_callback=ckpt_cb, limit_train_batches=2, limit_val_batches=1)
    w0 = model.model.unet.dec_l5.conv_block.conv2.conv.weight.detach().clone()
    trainer.fit(model, dm)
    ops.sync_side_streams()
    assert trainer.global_step == 1  # 3 tiles, batch 2, drop_last -> one step
    w1 = model.model.unet.dec_l5.conv_block.conv2.conv.weight.detach()
    assert not torch.equal(w0.to(w1.device), w1), "the optimizer step did not change the weights"
    assert "f1_score" in model.logged and "val_loss" in model.logged and np.isfinite(float(model.logged["val_loss"]))
    last = os.path.join(results, "checkpoints", "last.ckpt")
    assert os.path.exists(last)

    # eval: checkpoint round trip (strict keys), TTA forward, probabilities on disk as the reference writes them
    model2 = Model.load_from_checkpoint(last)
    sd1, sd2 = model.state_dict(), model2.state_dict()
    assert sd1.keys() == sd2.keys() and all(torch.equal(sd1[k].cpu(), sd2[k].cpu()) for k in sd1)
    os.makedirs(os.path.join(results, "probs"))
    os.makedirs(os.path.join(results, "targets"))
    Trainer(gpus=1, default_root_dir=results).test(model2, test_dataloaders=dm.test_dataloader())
    probs = sorted(os.listdir(os.path.join(results, "probs")))
    assert len(probs) == 3 and probs[0].startswith("test_localization_00000" if task == "pre" else "test_damage_00000")
    p0 = np.load(os.path.join(results, "probs", probs[0]))
    assert p0.dtype == np.float32 and 0.0 <= p0.min() and p0.max() <= 1.0
    assert "f1_score" in model2.logged
    if task == "post":  # softmax over the 4 damage classes, planar like the reference's np.save (plt.py:129-131)
        assert p0.shape == (4, 1024, 1024) and np.allclose(p0.sum(0), 1.0, atol=1e-4)
        return
    assert p0.shape == (1024, 1024)

    # post-process from the stored probabilities (loc: (h, w); dmg would be (4, h, w)) == numpy restatement, bit for bit
    loc = torch.from_numpy(p0).cuda()
    dmg = torch.rand(4, 1024, 1024, generator=torch.Generator().manual_seed(3)).cuda()
    pre_map, post_map = ops.post_process_probs(loc, dmg)
    post_ref = np.argmax(dmg.cpu().numpy(), axis=0) + 1
    pre_ref = (p0 > 0.3) | ((p0 > 0.1) & (post_ref > 1))
    assert np.array_equal(pre_map.cpu().numpy().astype(bool), pre_ref)
    assert np.array_equal(post_map.cpu().numpy(), (post_ref * pre_ref).astype(np.uint8))


def test_post_process_cli(tmp_path):
    """xview2_b200.utils.post_process (reference utils/post_process.py): probs/*.npy -> predictions/*.png, against a numpy
    restatement of post_process.py:27-47 including the connected-component vote and the dilation."""
    from PIL import Image
    from scipy.ndimage import grey_dilation, label

    from xview2_b200.utils import post_process as pp

    rng = np.random.default_rng(11)
    os.makedirs(tmp_path / "probs")
    cases = []
    for i in range(2):
        coarse = rng.random((64, 64)).astype(np.float32)
        loc = np.kron(coarse, np.ones((16, 16), np.float32)) * 0.9 + rng.random((1024, 1024)).astype(np.float32) * 0.1
        dmg = rng.random((4, 1024, 1024)).astype(np.float32)
        dmg /= dmg.sum(0, keepdims=True)
        np.save(tmp_path / "probs" / f"test_localization_{i:05d}.npy", loc)
        np.save(tmp_path / "probs" / f"test_damage_{i:05d}.npy", dmg)
        cases.append((loc, dmg))
    assert pp.main(["--results", str(tmp_path), "--components", "--dilate", "--dilation_rate", "3"]) == 2
    for i, (loc, dmg) in enumerate(cases):
        post = np.argmax(dmg, axis=0) + 1
        pre = ((loc > 0.3) | ((loc > 0.1) & (post > 1))).astype(np.int64)
        post = post * pre
        components, n = label(post > 0)
        for b in range(1, n + 1):
            labels, counts = np.unique(post[components == b], return_counts=True)
            post[components == b] = labels[np.argmax(counts)]
        pre, post = grey_dilation(pre, size=(3, 3)), grey_dilation(post, size=(3, 3))
        got_pre = np.array(Image.open(tmp_path / "predictions" / f"test_localization_{i:05d}_prediction.png"))
        got_post = np.array(Image.open(tmp_path / "predictions" / f"test_damage_{i:05d}_prediction.png"))
        assert np.array_equal(got_pre, pre.astype(np.uint8)) and np.array_equal(got_post, post.astype(np.uint8))
