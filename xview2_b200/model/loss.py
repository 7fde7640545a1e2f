"""Loss with the reference's interface (/root/reference/model/loss.py): ``Loss(args)(y_pred, y_true)`` and the
module-level ``losses`` registry.  dice / focal / ce / ohem are computed by ONE fused reduction pass + ONE backward pass
(xview2_b200/csrc/loss.cu) instead of MONAI's chain of full-resolution ATen ops; the `post` masking (loss.py:86-90) is
done in-kernel, so no boolean-mask gather and no host synchronisation happen.

'ohem' reproduces what the reference actually computes: loss.py:45 slices the (values, indices) tuple returned by
sort(), so no negative is ever discarded and the result equals the mean cross-entropy.
'mse' and 'coral' (ordinal damage heads, unet.py:21-26, loss.py:54-65,92-94) run as their own reduction + backward kernels
(xview2_b200/csrc/resample.cu); like the reference they are used on their own, not '+'-joined with other terms.
"""
from torch import nn

from .. import ops


class _Term(nn.Module):
    def __init__(self, name):
        super().__init__()
        self.name = name

    def forward(self, y_pred, y_true, post=False):
        return ops.seg_loss(y_pred, y_true, self.name, post)


losses = {name: _Term(name) for name in ("dice", "focal", "ce", "ohem", "mse", "coral")}


class Loss(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.loss_str = args.loss_str
        self.post = args.type == "post"
        names = self.loss_str.split("+")
        for name in names:
            if name not in losses:
                raise NotImplementedError(f"loss '{name}' is not one of dice, focal, ce, ohem, mse, coral")
        if len(names) > 1 and any(n in ("mse", "coral") for n in names):
            raise NotImplementedError("mse / coral change the head (unet.py:21-26) and cannot be '+'-joined with other terms")
        self.losses = nn.ModuleList([losses[name] for name in self.loss_str.split("+")])

    def forward(self, y_pred, y_true, weight=1.0, label_stride=1):
        """`weight` / `label_stride` let Model.compute_loss fold the deep-supervision scale and the nearest label
        down-sampling (plt.py:73) into the kernel; with the defaults this is exactly Loss.forward."""
        return ops.seg_loss(y_pred, y_true, self.loss_str, self.post, weight, label_stride)
