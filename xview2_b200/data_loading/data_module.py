"""``DataModule(args)`` with the reference's surface (/root/reference/data_loading/data_module.py:8-40): three loader factories
over the xBD directory layout ``<data>/{train,test,holdout}/{images,targets}`` -- ``test`` is the validation split and
``holdout`` the test split, as in the reference."""
import os

from .pytorch_loader import fetch_pytorch_loader

_SPLIT_DIR = {"train": "train", "val": "test", "test": "holdout"}


class DataModule:
    def __init__(self, args):
        self.args = args
        for split, sub in _SPLIT_DIR.items():
            setattr(self, f"{split}_path", os.path.join(args.data, sub))
        common = {"pin_memory": True, "num_workers": args.num_workers, "seed": getattr(args, "seed", 1)}
        # training: shuffled, whole batches only; evaluation: file order, ragged last batch kept (data_module.py:16-29)
        self.train_loader_kwargs = dict(common, batch_size=args.batch_size, shuffle=True, drop_last=True)
        self.test_loader_kwargs = dict(common, batch_size=args.val_batch_size, shuffle=False, drop_last=False)

    def _loader(self, split):
        training = split == "train"
        kwargs = self.train_loader_kwargs if training else self.test_loader_kwargs
        return fetch_pytorch_loader(getattr(self, f"{split}_path"), self.args.type, training, kwargs,
                                    self.args.autoaugment if training else False)

    def train_dataloader(self):
        return self._loader("train")

    def val_dataloader(self):
        return self._loader("val")

    def test_dataloader(self):
        return self._loader("test")
