"""Minimal stand-in for the pytorch_lightning 1.0 ``Trainer`` the reference configures at /root/reference/main.py:96-122,
driving the same ``Model`` hooks in the same order (fit: training_step -> backward -> optimizer/scheduler step per batch,
then on_validation_epoch_start / validation_step* / validation_epoch_end per epoch; test: on_test_epoch_start / test_step* /
test_epoch_end), with ``ModelCheckpoint`` / ``EarlyStopping`` on the logged ``f1_score``.

Data parallelism is the B200-native scheme of SURVEY.md 8(e): ONE PROCESS PER GPU (``--gpus N`` re-executes the script under
``torch.distributed.run`` exactly like PL's ddp accelerator re-executes it with LOCAL_RANK set), tiles sharded over ranks by
the loader, and a single NCCL all-reduce of the flat gradient buffer per step (xview2_b200.optim.FlatParams); BN statistics
stay rank-local (``sync_batchnorm`` is accepted and ignored: north_star keeps the gradient all-reduce as the only collective).
"""
import os
import subprocess
import sys

import torch

from . import lib


def seed_everything(seed):
    import random

    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    return seed


class EarlyStopping:
    def __init__(self, monitor="f1_score", patience=100, verbose=False, mode="max"):
        self.monitor, self.patience, self.verbose, self.mode = monitor, patience, verbose, mode
        self.best, self.wait = None, 0

    def should_stop(self, logged):
        if self.monitor not in logged:
            return False
        v = float(logged[self.monitor])
        better = self.best is None or (v > self.best if self.mode == "max" else v < self.best)
        if better:
            self.best, self.wait = v, 0
            return False
        self.wait += 1
        return self.wait >= self.patience


class ModelCheckpoint:
    def __init__(self, monitor="f1_score", mode="max", save_last=True, dirpath=None):
        self.monitor, self.mode, self.save_last, self.dirpath = monitor, mode, save_last, dirpath
        self.best, self.best_model_path = None, None

    def on_epoch_end(self, trainer, model, optimizer):
        if trainer.global_rank != 0:
            return
        d = self.dirpath or os.path.join(trainer.default_root_dir, "checkpoints")
        os.makedirs(d, exist_ok=True)
        ckpt = model.checkpoint()
        ckpt["optimizer"] = optimizer.state_dict() if optimizer is not None else None
        ckpt["global_step"] = trainer.global_step
        sched = getattr(trainer, "scheduler", None)
        ckpt["lr_schedulers"] = [sched.state_dict()] if sched is not None else []  # PL's checkpoint key
        if self.save_last:
            torch.save(ckpt, os.path.join(d, "last.ckpt"))
        if self.monitor in model.logged:
            v = float(model.logged[self.monitor])
            if self.best is None or (v > self.best if self.mode == "max" else v < self.best):
                self.best = v
                if self.best_model_path and os.path.exists(self.best_model_path):
                    os.remove(self.best_model_path)
                self.best_model_path = os.path.join(d, f"epoch={model.current_epoch}.ckpt")
                torch.save(ckpt, self.best_model_path)


class Trainer:
    def __init__(self, gpus=1, logger=False, precision=16, benchmark=True, deterministic=False, num_sanity_val_steps=0,
                 callbacks=None, max_epochs=1, min_epochs=1, sync_batchnorm=False, accelerator=None, default_root_dir=".",
                 checkpoint_callback=None, resume_from_checkpoint=None, limit_train_batches=None, limit_val_batches=None,
                 use_cuda_graph=False):
        self.gpus = max(1, int(gpus))
        self.precision = precision
        self.callbacks = callbacks or []
        self.max_epochs, self.min_epochs = max_epochs, min_epochs
        # main.py:106 passes sync_batchnorm=gpus>1; north_star keeps ONE collective (the gradient all-reduce), so BN
        # statistics are rank-local here and the flag is accepted for signature parity only.
        self.sync_batchnorm = False
        if sync_batchnorm:
            import warnings
            warnings.warn("sync_batchnorm=True is accepted for signature parity but BatchNorm statistics stay rank-local here "
                          "(the gradient all-reduce is the only collective of this path); at very small per-GPU batches "
                          "(e.g. 2 pairs) this differs from the reference's SyncBatchNorm run", stacklevel=2)
        self.default_root_dir = default_root_dir
        self.checkpoint_callback = checkpoint_callback
        self.resume_from_checkpoint = resume_from_checkpoint
        self.limit_train_batches, self.limit_val_batches = limit_train_batches, limit_val_batches
        self.use_cuda_graph = use_cuda_graph  # opt-in: replay forward + backward from one CUDA graph (xview2_b200.graph)
        self.datamodule = None
        self.global_rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.global_step = 0
        self.scheduler = None

    # -- process-per-GPU launch ----------------------------------------------------------------------------------
    def _maybe_spawn(self):
        """``--gpus N`` without a launcher: re-execute this script under torch.distributed.run, one rank per GPU."""
        if self.gpus <= 1 or "WORLD_SIZE" in os.environ:
            return
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={self.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"), *sys.argv]
        sys.exit(subprocess.call(cmd))

    def _setup(self, model):
        self._maybe_spawn()
        if not torch.cuda.is_available():
            raise lib.Xv2Error("xview2_b200 runs on CUDA devices only (there is no CPU fallback)")
        torch.cuda.set_device(self.local_rank)
        lib.init(self.local_rank)
        if self.world_size > 1 and not torch.distributed.is_initialized():
            torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        model.trainer = self
        return model.to(torch.device("cuda", self.local_rank))

    @staticmethod
    def _limited(loader, limit):
        for i, batch in enumerate(loader):
            if limit is not None and i >= limit:
                break
            yield i, batch

    # -- fit ---------------------------------------------------------------------------------------------------
    def fit(self, model, datamodule):
        self.datamodule = datamodule
        model = self._setup(model)
        start_epoch = 0
        resume = None
        if self.resume_from_checkpoint and os.path.exists(self.resume_from_checkpoint):
            resume = torch.load(self.resume_from_checkpoint, map_location="cpu", weights_only=False)
            model.load_reference_state_dict(resume["state_dict"])
            start_epoch = int(resume.get("epoch", -1)) + 1
        train_loader = datamodule.train_dataloader()
        val_loader = datamodule.val_dataloader()
        conf = model.configure_optimizers()
        optimizer, scheduler = (conf["optimizer"], conf["lr_scheduler"]["scheduler"]) if isinstance(conf, dict) else (conf, None)
        flat = model.flat
        flat.broadcast_params(0)
        if self.world_size > 1:
            flat.enable_bucketed_allreduce(model)  # DDP-style overlap of the gradient all-reduce with backward (main.py:106-107)
        from . import ops  # weight gradients stay on the compute stream (a side stream measured neutral: DESIGN.md section 4)
        self.scheduler = scheduler
        if resume is not None and resume.get("optimizer"):
            optimizer.load_state_dict(resume["optimizer"])
        if resume is not None:
            # PL restores global_step and the lr_schedulers states; without them the Noam warm-up would be replayed
            self.global_step = int(resume.get("global_step", start_epoch * max(1, len(train_loader))))
            if scheduler is not None:
                states = resume.get("lr_schedulers") or []
                if states:
                    scheduler.load_state_dict(states[0])
                else:  # checkpoint without scheduler state: re-derive the position from the step count
                    scheduler.step(self.global_step + 1)
        graphed = None
        # all training work runs on one non-default stream, so that a CUDA-graph capture can share it with the eager steps
        train_stream = torch.cuda.Stream() if self.use_cuda_graph else torch.cuda.current_stream()
        train_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(train_stream):
            self._fit_epochs(model, datamodule, train_loader, val_loader, optimizer, scheduler, flat, start_epoch, train_stream)
        torch.cuda.current_stream().wait_stream(train_stream)
        return model

    def _fit_epochs(self, model, datamodule, train_loader, val_loader, optimizer, scheduler, flat, start_epoch, train_stream):
        graphed = None
        for epoch in range(start_epoch, self.max_epochs):
            model.current_epoch = epoch
            model.train()
            if hasattr(train_loader, "set_epoch"):
                train_loader.set_epoch(epoch)
            for i, batch in self._limited(train_loader, self.limit_train_batches):
                if self.use_cuda_graph and graphed is None and self.global_step >= 1:
                    try:  # static shapes (drop_last loader): after one eager step (caches warm) capture forward + backward once
                        from .graph import GraphedTrainStep
                        graphed = GraphedTrainStep(model, optimizer, batch, warmup=0, stream=train_stream)
                    except Exception as exc:  # noqa: BLE001 -- capture is an optimisation; the eager step is always valid
                        print(f"CUDA-graph capture unavailable ({type(exc).__name__}: {exc}); running eagerly", flush=True)
                        self.use_cuda_graph = False
                if graphed is not None:
                    loss = graphed(batch)
                else:
                    optimizer.zero_grad()
                    loss = model.training_step(batch, i)
                    loss.backward()
                    optimizer.grad_scale = 1.0 / flat.all_reduce_grads()
                    optimizer.step()
                if scheduler is not None:
                    scheduler.step()
                self.global_step += 1
            model.eval()
            model.on_validation_epoch_start()
            outputs = []
            with torch.no_grad():
                for i, batch in self._limited(val_loader, self.limit_val_batches):
                    outputs.append(model.validation_step(batch, i))
            if outputs:
                model.validation_epoch_end(outputs)
            if self.checkpoint_callback is not None:
                self.checkpoint_callback.on_epoch_end(self, model, optimizer)
            if epoch + 1 >= self.min_epochs and any(cb.should_stop(model.logged) for cb in self.callbacks
                                                    if isinstance(cb, EarlyStopping)):
                break
        return model

    # -- test ----------------------------------------------------------------------------------------------------
    def test(self, model, test_dataloaders=None):
        model = self._setup(model)
        model.eval()
        model.on_test_epoch_start()
        with torch.no_grad():
            for i, batch in enumerate(test_dataloaders):
                model.test_step(batch, i)
        model.test_epoch_end(None)
        return [dict(model.logged)]
