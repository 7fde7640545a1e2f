"""Noam learning-rate schedule with the reference's constructor (/root/reference/utils/scheduler.py:6-59): linear warm-up
from init_lr to max_lr over warmup_epochs * steps_per_epoch steps, then exponential decay to final_lr at total steps.
Host-side scalar arithmetic; stepped once per optimizer step."""
import numpy as np


class NoamLR:
    def __init__(self, optimizer, warmup_epochs, total_epochs, steps_per_epoch, init_lr, max_lr, final_lr,
                 fine_tune_coff=1.0, fine_tune_param_idx=0):
        self.optimizer = optimizer
        self.num_lrs = len(optimizer.param_groups)
        n = self.num_lrs
        self.steps_per_epoch = steps_per_epoch
        self.init_lr = np.array([init_lr] * n, dtype=np.float64)
        self.max_lr = np.array([max_lr] * n, dtype=np.float64)
        self.final_lr = np.array([final_lr] * n, dtype=np.float64)
        self.lr_coff = np.array([1.0] * n)
        self.lr_coff[fine_tune_param_idx] = fine_tune_coff
        self.current_step = 0
        self.lr = [init_lr] * n
        self.warmup_steps = (np.array([warmup_epochs] * n) * steps_per_epoch).astype(int)
        self.total_steps = np.array([total_epochs] * n) * steps_per_epoch
        self.linear_increment = (self.max_lr - self.init_lr) / np.maximum(self.warmup_steps, 1)
        self.exponential_gamma = (self.final_lr / self.max_lr) ** (1 / np.maximum(self.total_steps - self.warmup_steps, 1))
        for i, group in enumerate(optimizer.param_groups):
            group["lr"] = self.lr[i]

    def get_lr(self):
        return list(self.lr)

    def step(self, current_step=None):
        self.current_step = current_step if current_step is not None else self.current_step + 1
        for i in range(self.num_lrs):
            if self.current_step <= self.warmup_steps[i]:
                lr = self.init_lr[i] + self.current_step * self.linear_increment[i]
            elif self.current_step <= self.total_steps[i]:
                lr = self.max_lr[i] * (self.exponential_gamma[i] ** (self.current_step - self.warmup_steps[i]))
            else:
                lr = self.final_lr[i]
            self.lr[i] = float(lr * self.lr_coff[i])
            self.optimizer.param_groups[i]["lr"] = self.lr[i]

    def state_dict(self):
        return {"current_step": self.current_step, "lr": list(self.lr)}

    def load_state_dict(self, sd):
        self.current_step = sd["current_step"]
        self.lr = list(sd["lr"])
        for i, group in enumerate(self.optimizer.param_groups):
            group["lr"] = self.lr[i]
