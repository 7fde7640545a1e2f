"""F1 metric with the reference's interface (/root/reference/utils/f1.py:18-56): ``F1(args).update(preds, targets)``,
``compute()``, ``reset()``.  softmax -> argmax -> tp/fp/fn counting is ONE libxv2 pass (xv2_f1_update) over the logits
that accumulates int64 counters on the device (softmax is monotone, so the argmax of the logits is taken directly;
ties -> lowest index like torch.argmax).  Cross-rank reduction is one sum all-reduce of 3*(n_class-1) counters at
compute() time (dist_reduce_fx="sum", f1.py:24-26).
"""
import torch

from .. import ops


def convert_to_labels(loss_str, logits):
    """f1.py:7-15: argmax + 1, or the ordinal decoding of the mse / coral heads (one libxv2 pass)."""
    if loss_str in ("mse", "coral"):
        return ops.ordinal_labels(logits, loss_str, want_u8=True).long()
    return torch.argmax(logits, dim=1) + 1


class F1:
    def __init__(self, args):
        self.loss_str = args.loss_str
        self.n_class = 2 if args.type == "pre" else 5
        self.counters = None  # int64 [3 * (n_class - 1)] = tp | fp | fn, created on the first update's device

    # -- state -----------------------------------------------------------------------------------------------
    def reset(self):
        if self.counters is not None:
            self.counters.zero_()

    @property
    def tp(self):
        return self._part(0)

    @property
    def fp(self):
        return self._part(1)

    @property
    def fn(self):
        return self._part(2)

    def _part(self, i):
        k = self.n_class - 1
        if self.counters is None:
            return torch.zeros(k)
        return self.counters[i * k:(i + 1) * k].float()

    def load_counts(self, tp, fp, fn):
        """Restores the tp / fp / fn states of a checkpoint (pytorch_lightning Metric states are persistent)."""
        c = torch.cat([torch.as_tensor(t).reshape(-1).to(torch.int64) for t in (tp, fp, fn)])
        if c.numel() != 3 * (self.n_class - 1):
            raise ValueError(f"F1 state of {c.numel()} counters does not fit n_class={self.n_class}")
        self.counters = c if self.counters is None else c.to(self.counters.device)

    # -- Metric API -------------------------------------------------------------------------------------------
    def update(self, preds, targets, pred_map=None):
        if self.counters is None:
            self.counters = torch.zeros(3 * (self.n_class - 1), dtype=torch.int64, device=preds.device)
        elif self.counters.device != preds.device:
            self.counters = self.counters.to(preds.device)
        if self.n_class == 5 and self.loss_str in ("mse", "coral"):
            got = ops.ordinal_labels(preds, self.loss_str, targets, self.counters, clamp4=True, want_u8=pred_map is not None)
            if pred_map is not None:
                pred_map.copy_(got)
            return
        ops.f1_update(preds, targets, self.n_class, self.counters, pred_map)

    def __call__(self, preds, targets):
        self.update(preds, targets)

    def _synced(self):
        c = self.counters
        if c is None:
            return torch.zeros(3 * (self.n_class - 1), dtype=torch.int64)
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            c = c.clone()
            torch.distributed.all_reduce(c)
        return c

    def compute(self):
        """f1.py:44-49: returns (f1, per_damage_class_f1 | None) as CPU tensors."""
        k = self.n_class - 1
        c = self._synced().cpu().float()
        tp, fp, fn = c[:k], c[k:2 * k], c[2 * k:]
        f1_score = 200 * tp / (2 * tp + fp + fn)
        if self.n_class == 5:
            f1 = 4 / sum((f1_ + 1e-6) ** -1 for f1_ in f1_score)
            return f1, f1_score
        return f1_score, None
