"""Command line of the reference (/root/reference/main.py:26-125): same flags, defaults and behaviour, driving the
B200-native ``Model`` / ``DataModule`` / ``Trainer`` of xview2_b200.

    python main.py --type pre --encoder resnest50 --data /data --results /results --gpus 8 --precision bf16
    python main.py --exec_mode eval --type pre --ckpt /results/checkpoints/last.ckpt --tta

Differences: ``--precision`` also accepts ``bf16`` (16 selects the same bf16 tensor-core path: fp16 + GradScaler has no
counterpart here); ``--gpus N`` re-executes the script under torch.distributed.run (one process per GPU).
"""
import os
import shutil
from argparse import ArgumentDefaultsHelpFormatter, ArgumentParser

import torch

from xview2_b200.data_loading.data_module import DataModule
from xview2_b200.model.plt import Model
from xview2_b200.trainer import EarlyStopping, ModelCheckpoint, Trainer, seed_everything
from xview2_b200.utils.gpu_affinity import set_affinity


def make_empty_dir(path):
    shutil.rmtree(path, ignore_errors=True)
    os.makedirs(path)


def set_cuda_devices(gpus):
    assert gpus <= torch.cuda.device_count(), f"Requested {gpus} gpus, available {torch.cuda.device_count()}."
    device_list = ",".join([str(i) for i in range(gpus)])
    os.environ["CUDA_VISIBLE_DEVICES"] = os.environ.get("CUDA_VISIBLE_DEVICES", device_list)


def _precision(v):
    return "bf16" if v == "bf16" else int(v)


# The reference's 15 top-level flags (main.py:29-53): (flag, type, default, choices, help).  Names and defaults are the
# contract; the help texts are ours.
_BASE_FLAGS = (
    ("exec_mode", str, "train", ["train", "eval"], "train a model, or evaluate a checkpoint on the hold-out split"),
    ("data", str, "/data", None, "dataset root holding train/, test/ (validation) and holdout/ (test)"),
    ("results", str, "/results", None, "output directory: checkpoints/, logs, probs/ and targets/"),
    ("gpus", int, 1, None, "B200s to use; > 1 starts one process per GPU"),
    ("num_workers", int, 8, None, "PNG decode threads per process"),
    ("batch_size", int, 16, None, "tiles (or pre/post pairs) per GPU and training step"),
    ("val_batch_size", int, 13, None, "tiles per GPU and evaluation step"),
    ("precision", _precision, 16, [16, 32, "bf16"], "16 / bf16: bf16 tensor-core kernels; 32: fp32 parity kernels"),
    ("epochs", int, 250, None, "number of training epochs"),
    ("patience", int, 100, None, "epochs without an F1 improvement before training stops early"),
    ("ckpt", str, None, None, "checkpoint to resume from (train) or to evaluate (eval)"),
    ("logname", str, "logs", None, "basename of the JSON-lines log in --results"),
    ("ckpt_pre", str, None, None, "localisation checkpoint whose encoder initialises the damage model (--type post)"),
    ("type", str, None, ["pre", "post"], "pre: building localisation; post: damage assessment on pre/post pairs"),
    ("seed", int, 1, None, "random seed"),
)


def build_parser():
    parser = ArgumentParser(formatter_class=ArgumentDefaultsHelpFormatter)
    for name, kind, default, choices, text in _BASE_FLAGS:
        extra = {"choices": choices} if choices else {}
        parser.add_argument(f"--{name}", type=kind, default=default, help=text, **extra)
    return Model.add_model_specific_args(parser)


def transplant_encoder(model, pretrained_state, dmg_model):
    """main.py:76-94: copy the localisation model's encoder tensors (names containing "enc") into the damage model.
    The reference's ``parallel`` branch indexes state_dict() with the whole key list (main.py:87) and fails; here the
    intended key2 is used."""
    own = model.state_dict()
    copied = 0
    for name, tensor in pretrained_state.items():
        if "enc" not in name:
            continue
        if "parallel" in dmg_model:
            targets = [name.replace("unet", "unet_pre"), name.replace("unet", "unet_post")]
        elif dmg_model == "siameseEnc":
            targets = [name.replace(".unet", "")]
        else:
            targets = [name]
        for key in targets:
            if key in own:
                own[key].copy_(tensor)
                copied += 1
    return copied


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.interpolate:
        args.deep_supervision = False
        args.dec_interp = False

    set_cuda_devices(args.gpus)
    set_affinity(os.getenv("LOCAL_RANK", "0"), "socket_unique_interleaved")
    seed_everything(args.seed)
    data_module = DataModule(args)

    callbacks = None
    model_ckpt = None
    checkpoint = args.ckpt if args.ckpt is not None and os.path.exists(args.ckpt) else None
    if args.exec_mode == "train":
        model = Model(args)
        model_ckpt = ModelCheckpoint(monitor="f1_score", mode="max", save_last=True)
        callbacks = [EarlyStopping(monitor="f1_score", patience=args.patience, verbose=True, mode="max")]
    else:
        assert args.ckpt is not None, "No checkpoint found for evaluation"
        model = Model.load_from_checkpoint(args.ckpt)
        model.args.results = args.results
        model.args.tta = args.tta or model.args.tta

    if args.type == "post" and args.ckpt_pre is not None:
        pretrained = torch.load(args.ckpt_pre, map_location="cpu", weights_only=False)["state_dict"]
        transplant_encoder(model, pretrained, args.dmg_model)

    trainer = Trainer(gpus=args.gpus, logger=False, precision=args.precision, benchmark=True, deterministic=False,
                      num_sanity_val_steps=0, callbacks=callbacks, max_epochs=args.epochs, min_epochs=args.epochs,
                      sync_batchnorm=args.gpus > 1, accelerator="ddp" if args.gpus > 1 else None,
                      default_root_dir=args.results, checkpoint_callback=model_ckpt, resume_from_checkpoint=checkpoint)

    if args.exec_mode == "train":
        trainer.fit(model, data_module)
    else:
        pred_dir = os.path.join(args.results, "probs")
        targets_dir = os.path.join(args.results, "targets")
        if not os.path.exists(pred_dir):
            make_empty_dir(pred_dir)
        if not os.path.exists(targets_dir):
            make_empty_dir(targets_dir)
        trainer.test(model, test_dataloaders=data_module.test_dataloader())


if __name__ == "__main__":
    main()
