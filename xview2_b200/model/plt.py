"""``Model`` with the reference's Lightning hooks (/root/reference/model/plt.py:20-234): same constructor, hook names,
argument meaning and CLI flags, so ``main.py`` and a user's training script read the same.  pytorch_lightning, apex,
dllogger and torch_optimizer are not dependencies: the hooks are driven by xview2_b200.trainer.Trainer, optimizers are
the fused flat-buffer ones of xview2_b200.optim, the dllogger JSON-lines shape is written directly.

Every arithmetic step of forward / loss / metric / post-process runs in libxv2 (hand-written sm_100a CUDA):
  forward           U-Net kernels; --tta folds the three flips and the 4-way logit mean into two kernels (plt.py:42-48)
  compute_loss      fused dice / focal / ce reduction + analytic backward, deep-supervision weights and the nearest
                    label down-sampling (plt.py:69-77) folded into the kernel arguments
  F1                one counting pass over logits (utils/f1.py)
  save / post-process  sigmoid / softmax -> .npy as the reference writes them (plt.py:126-144)
"""
import json
import os
from argparse import ArgumentParser

import numpy as np
import torch
from torch import nn

from .. import ops
from ..optim import FlatParams, FusedAdamW, FusedSGD
from ..utils.f1 import F1
from ..utils.scheduler import NoamLR
from .loss import Loss
from .unet import UNetLoc, get_dmg_unet


def compute_loss(loss_fn, preds, label, deep_supervision):
    """Model.compute_loss (plt.py:69-77).  The 1, 1/2, 1/4 weights, the 1/(2 - 2^-3) normalisation and the nearest
    label down-sampling are folded into the loss kernels (weight / label stride arguments)."""
    if not deep_supervision or not isinstance(preds, (list, tuple)):
        return loss_fn(preds, label)
    c_norm = 1 / (2 - 2 ** (-len(preds)))
    loss = loss_fn(preds[0], label, weight=c_norm)
    for i, pred in enumerate(preds[1:]):
        stride = label.shape[-1] // pred.shape[-1]
        loss = loss + loss_fn(pred, label, weight=c_norm * 0.5 ** (i + 1), label_stride=stride)
    return loss


class _JsonLinesLogger:
    """The two dllogger backends the reference configures (plt.py:35-40): a JSON-lines file and stdout."""

    def __init__(self, path):
        self.path = path
        self._fh = None

    def log(self, step, data):
        line = {"type": "LOG", "step": step if step != () else [], "data": data}
        if self.path is not None:
            if self._fh is None:
                os.makedirs(os.path.dirname(self.path) or ".", exist_ok=True)
                self._fh = open(self.path, "a")
            self._fh.write("DLLL " + json.dumps(line) + "\n")
        prefix = f"Epoch: {step} " if step != () else ""
        print(prefix + " ".join(f"{k}: {v}" for k, v in data.items()), flush=True)

    def flush(self):
        if self._fh is not None:
            self._fh.flush()


class Model(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.save_hyperparameters(args)
        self.args = args
        self.f1_score = F1(args)
        self.model = UNetLoc(args) if args.type == "pre" else get_dmg_unet(args)
        self.loss = Loss(args)
        self.best_f1 = torch.tensor(0)
        self.best_epoch = 0
        self.tta_flips = [[2], [3], [2, 3]]
        self.lr = args.lr
        self.n_class = 2 if self.args.type == "pre" else 5
        self.test_idx = 0
        results = getattr(args, "results", None)
        logname = getattr(args, "logname", "logs")
        self.dllogger = _JsonLinesLogger(os.path.join(results, f"{logname}.json") if results else None)
        # attributes Lightning would provide
        self.current_epoch = 0
        self.trainer = None
        self.logged = {}
        self.flat = None

    # -- Lightning-provided surface ----------------------------------------------------------------------------
    def save_hyperparameters(self, args):
        self.hparams = {"args": args}

    def log(self, name, value):
        self.logged[name] = value

    def train_dataloader(self):
        if self.trainer is None or self.trainer.datamodule is None:
            raise RuntimeError("no data module attached")
        return self.trainer.datamodule.train_dataloader()

    @classmethod
    def load_from_checkpoint(cls, path, map_location="cpu"):
        """main.py:74.  Reads a Lightning-style checkpoint dict: {"state_dict", "hyper_parameters": {"args": ...}}."""
        ckpt = torch.load(path, map_location=map_location, weights_only=False)
        model = cls(ckpt["hyper_parameters"]["args"])
        model.load_reference_state_dict(ckpt["state_dict"])
        return model

    def load_reference_state_dict(self, state_dict):
        """Strict load of a checkpoint written by the reference or by this package.  pytorch_lightning 1.0 Metric states are
        persistent, so reference checkpoints carry ``f1_score.tp / fp / fn`` (utils/f1.py:24-26); they are absorbed into the
        device counters of xview2_b200.utils.f1.F1 instead of tripping the strict key check."""
        sd = dict(state_dict)
        metric = {k[len("f1_score."):]: sd.pop(k) for k in list(sd) if k.startswith("f1_score.")}
        self.load_state_dict(sd, strict=True)
        if all(k in metric for k in ("tp", "fp", "fn")):
            self.f1_score.load_counts(metric["tp"], metric["fp"], metric["fn"])
        return self

    def checkpoint(self):
        sd = {k: v.detach().cpu().clone() for k, v in self.state_dict().items()}
        sd.update({f"f1_score.{k}": getattr(self.f1_score, k).detach().cpu().clone() for k in ("tp", "fp", "fn")})  # PL Metric states
        return {"state_dict": sd, "hyper_parameters": self.hparams, "epoch": self.current_epoch}

    # -- hooks (plt.py:42-67) -----------------------------------------------------------------------------------
    @staticmethod
    def _image(batch):
        """``batch["image"]`` as the reference's loader yields it (f32 B x C x H x W, already normalised), or the
        decoded uint8 HWC tiles of the native loader (``batch["tiles"]``, optional ``batch["tiles_post"]``), which are
        normalised + laid out NHWC bf16 by one kernel (pytorch_loader.py:63,169-170)."""
        if "image" in batch:
            return batch["image"]
        return ops.normalize_tiles(batch["tiles"], batch.get("tiles_post"), torch.bfloat16)

    def forward(self, img):
        pred = self.model(img)
        if self.args.tta:
            flipped = [ops.flip(self.model(ops.flip(img, dims)), dims) for dims in self.tta_flips]
            pred = ops.mean4(pred, *flipped)
        return pred

    def training_step(self, batch, _):
        img, lbl = self._image(batch), batch["mask"]
        pred = self.model(img)
        return self.compute_loss(pred, lbl)

    def validation_step(self, batch, _):
        img, lbl = self._image(batch), batch["mask"]
        pred = self.forward(img)
        loss = self.loss(pred, lbl)
        self.f1_score.update(pred, lbl)
        return {"val_loss": loss}

    def test_step(self, batch, batch_idx):
        img, lbl = self._image(batch), batch["mask"]
        pred = self.forward(img)
        self.f1_score.update(pred, lbl)
        self.save(pred, lbl)

    def compute_loss(self, preds, label):
        return compute_loss(self.loss, preds, label, self.args.deep_supervision)

    @staticmethod
    def metric_mean(name, outputs):
        return torch.stack([out[name] for out in outputs]).mean(dim=0)

    @staticmethod
    def update_damage_scores(metrics, dmgs_f1):
        if dmgs_f1 is not None:
            for i in range(4):
                metrics.update({f"D{i+1}": round(dmgs_f1[i].item(), 3)})

    def on_validation_epoch_start(self):
        self.f1_score.reset()

    def on_test_epoch_start(self):
        self.f1_score.reset()

    def validation_epoch_end(self, outputs):
        loss = self.metric_mean("val_loss", outputs)
        f1_score, dmgs_f1 = self.f1_score.compute()
        self.f1_score.reset()
        if self.n_class == 2:
            f1_score = f1_score.reshape(())
        if f1_score >= self.best_f1:
            self.best_f1 = f1_score
            self.best_epoch = self.current_epoch
        if int(os.getenv("LOCAL_RANK", "0")) == 0:
            metrics = {"f1": round(f1_score.item(), 3), "val_loss": round(loss.item(), 3),
                       "top_f1": round(self.best_f1.item(), 3)}
            self.update_damage_scores(metrics, dmgs_f1)
            self.dllogger.log(step=self.current_epoch, data=metrics)
            self.dllogger.flush()
        self.log("f1_score", f1_score.cpu())
        self.log("val_loss", loss.cpu())

    def test_epoch_end(self, _):
        f1_score, dmgs_f1 = self.f1_score.compute()
        self.f1_score.reset()
        if self.n_class == 2:
            f1_score = f1_score.reshape(())
        if int(os.getenv("LOCAL_RANK", "0")) == 0:
            metrics = {"f1": round(f1_score.item(), 3)}
            self.update_damage_scores(metrics, dmgs_f1)
            self.dllogger.log(step=(), data=metrics)
            self.dllogger.flush()
        self.log("f1_score", f1_score.cpu())

    def save(self, preds, targets):
        """plt.py:126-144: sigmoid(pred[:, 1]) | softmax(pred) -> <results>/probs/*.npy, targets -> PNG."""
        if self.args.type != "pre" and self.args.loss_str == "coral":    # plt.py:129: sum(sigmoid > 0.5) + 1, int64
            probs = ops.ordinal_labels(preds, "coral", want_u8=True).long().cpu().numpy()
        elif self.args.type != "pre" and self.args.loss_str == "mse":    # plt.py:131: round(relu(x0)) + 1, float32, unclamped
            probs = ops.ordinal_labels(preds, "mse", clamp4=False, want_f32=True).cpu().numpy()
        else:
            probs = ops.save_probs(preds).cpu().numpy()
        targets = targets.cpu().numpy().astype(np.uint8)
        from PIL import Image
        for prob, target in zip(probs, targets):
            task = "localization" if self.args.type == "pre" else "damage"
            fname = os.path.join(self.args.results, "probs", f"test_{task}_{self.test_idx:05d}")
            self.test_idx += 1
            np.save(fname, prob)
            Image.fromarray(target).save(fname.replace("probs", "targets") + "_target.png")

    @staticmethod
    def flip(data, axis):
        return ops.flip(data, axis)

    def configure_optimizers(self):
        """plt.py:150-179.  Parameters are re-homed into one flat fp32 buffer (xview2_b200.optim.FlatParams) so that the
        step is ONE kernel and the data-parallel exchange ONE all-reduce."""
        if self.flat is None:
            self.flat = FlatParams(self)
        name = self.args.optimizer.lower()
        if name in ("adamw", "adam"):  # apex FusedAdam defaults to adam_w_mode=True: both are decoupled-decay Adam
            optimizer = FusedAdamW(self.flat, lr=self.lr, weight_decay=self.args.weight_decay)
        elif name == "sgd":
            optimizer = FusedSGD(self.flat, lr=self.lr, momentum=self.args.momentum)
        else:
            raise NotImplementedError(f"optimizer '{name}' is outside the accelerated path (adamw, adam, sgd)")
        if not self.args.use_scheduler:
            return optimizer
        scheduler = {
            "scheduler": NoamLR(optimizer=optimizer, warmup_epochs=self.args.warmup, total_epochs=self.args.epochs,
                                # plt.py:170 divides an UNSHARDED DataLoader length by args.gpus; TileLoader.__len__ is already
                                # the per-rank batch count (shard_indices uses WORLD_SIZE), so it is used as is
                                steps_per_epoch=len(self.train_dataloader()),
                                init_lr=self.args.init_lr, max_lr=self.args.lr, final_lr=self.args.final_lr),
            "interval": "step",
            "frequency": 1,
        }
        return {"optimizer": optimizer, "lr_scheduler": scheduler}

    # The reference's 21 model flags (plt.py:185-233): names, defaults and choices are the contract, help texts are ours.
    _VALUE_FLAGS = (
        ("optimizer", str, "adamw", ["sgd", "adam", "adamw", "radam", "adabelief", "adabound", "adamp", "novograd"],
         "optimizer; adamw / adam / sgd run as fused flat-buffer kernels"),
        ("dmg_model", str, "siamese", ["siamese", "siameseEnc", "fused", "fusedEnc", "parallel", "parallelEnc", "diff", "cat"],
         "how the pre and post images are combined for damage assessment"),
        ("encoder", str, "resnest200", ["resnest50", "resnest101", "resnest200", "resnest269", "resnet50", "resnet101", "resnet152"],
         "encoder of the U-Net"),
        ("loss_str", str, "focal+dice", None, "'+'-joined loss terms out of dice, focal, ce, ohem; or mse / coral alone (ordinal damage heads)"),
        ("warmup", int, 1, None, "Noam schedule: warm-up epochs"),
        ("init_lr", float, 1e-4, None, "Noam schedule: learning rate at step 0"),
        ("final_lr", float, 1e-4, None, "Noam schedule: learning rate at the last step"),
        ("lr", float, 3e-4, None, "learning rate (the peak when the Noam schedule is on)"),
        ("weight_decay", float, 0, None, "decoupled weight decay"),
        ("momentum", float, 0.9, None, "SGD momentum"),
        ("dilation", int, 1, [1, 2, 4], "2 / 4: replace the stride of the last one / two encoder stages by dilation"),
    )
    _SWITCHES = (
        ("use_scheduler", "step the Noam learning-rate schedule"),
        ("tta", "average the logits over the four flips at evaluation"),
        ("ppm", "pyramid pooling module on the last encoder stage"),
        ("aspp", "atrous spatial pyramid pooling on the last encoder stage"),
        ("no_skip", "decoder without skip connections"),
        ("deep_supervision", "auxiliary heads on the two coarser decoder stages"),
        ("attention", "attention gates on the skip connections"),
        ("autoaugment", "ImageNet auto-augment policy (not accelerated)"),
        ("interpolate", "bilinear head on the encoder output instead of a decoder"),
        ("dec_interp", "3x3 conv + bilinear up-sampling instead of the transposed conv in the decoder"),
    )

    @classmethod
    def add_model_specific_args(cls, parent_parser):
        parser = ArgumentParser(parents=[parent_parser], add_help=False)
        for name, kind, default, choices, text in cls._VALUE_FLAGS:
            extra = {"choices": choices} if choices else {}
            parser.add_argument(f"--{name}", type=kind, default=default, help=text, **extra)
        for name, text in cls._SWITCHES:
            parser.add_argument(f"--{name}", action="store_true", help=text)
        return parser
