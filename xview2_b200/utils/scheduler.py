"""Noam learning-rate schedule behind the reference's ``NoamLR`` constructor (/root/reference/utils/scheduler.py:6-59).

The schedule is a closed form in the step index t (stepped once per optimizer step, plt.py:163-178):

    t <= W      : lr = init + t * (peak - init) / W                   linear warm-up over W = warmup_epochs * steps_per_epoch
    W < t <= T  : lr = peak * (final / peak) ** ((t - W) / (T - W))   exponential decay down to `final` at T = total steps
    t > T       : lr = final

times a per-group coefficient (``fine_tune_coff`` for group ``fine_tune_param_idx``, 1 elsewhere).  Host-side scalar math.
"""


def noam_value(t, warmup_steps, total_steps, init_lr, peak_lr, final_lr):
    if t <= warmup_steps:
        return init_lr + t * (peak_lr - init_lr) / max(warmup_steps, 1)
    if t <= total_steps:
        return peak_lr * (final_lr / peak_lr) ** ((t - warmup_steps) / max(total_steps - warmup_steps, 1))
    return final_lr


class NoamLR:
    def __init__(self, optimizer, warmup_epochs, total_epochs, steps_per_epoch, init_lr, max_lr, final_lr,
                 fine_tune_coff=1.0, fine_tune_param_idx=0):
        self.optimizer = optimizer
        groups = optimizer.param_groups
        self.steps_per_epoch = steps_per_epoch
        self.warmup_steps = int(warmup_epochs * steps_per_epoch)
        self.total_steps = total_epochs * steps_per_epoch
        self.shape = (float(init_lr), float(max_lr), float(final_lr))
        self.coefficients = [fine_tune_coff if g == fine_tune_param_idx else 1.0 for g in range(len(groups))]
        self.current_step = 0
        self.lr = [init_lr for _ in groups]
        self._publish()
        # torch's _LRScheduler.__init__ (the reference's base class, scheduler.py:41) calls self.step() once at construction,
        # so the reference trains its first batch at the step-1 learning rate and stays one step ahead throughout
        self.step()

    def _publish(self):
        for group, lr in zip(self.optimizer.param_groups, self.lr):
            group["lr"] = lr

    def get_lr(self):
        return list(self.lr)

    def step(self, current_step=None):
        self.current_step = self.current_step + 1 if current_step is None else current_step
        base = noam_value(self.current_step, self.warmup_steps, self.total_steps, *self.shape)
        self.lr = [float(base * k) for k in self.coefficients]
        self._publish()

    def state_dict(self):
        return {"current_step": self.current_step, "lr": list(self.lr)}

    def load_state_dict(self, sd):
        self.current_step = sd["current_step"]
        self.lr = list(sd["lr"])
        self._publish()
