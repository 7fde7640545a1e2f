"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Stand-in for ``monai.losses`` (monai==0.4.0, /root/reference/requirements.txt:12), which is
not installed here.  Only the two classes and the constructor arguments the reference uses
(/root/reference/model/loss.py:11-13) are restated, from the published MONAI 0.4.0 algorithm:

* DiceLoss(include_background, softmax=True, to_onehot_y=True, batch=True):
  softmax over channels; one-hot the index target; optionally drop channel 0 of both;
  sum over spatial dims AND batch (batch=True); f = 1 - (2*I + 1e-5) / (G + P + 1e-5); mean over channels.
* FocalLoss(gamma=2.0): logpt = log_softmax(x).gather(target); loss = mean(-(1-exp(logpt))**gamma * logpt).

Parity is UNPINNED against real MONAI (not installable offline); it is pinned only against
torch-native formulations in tests/test_oracle.py.
"""
import torch
import torch.nn.functional as F
from torch import nn


class DiceLoss(nn.Module):
    def __init__(self, include_background=True, to_onehot_y=False, sigmoid=False, softmax=False,
                 squared_pred=False, jaccard=False, reduction="mean", smooth_nr=1e-5, smooth_dr=1e-5, batch=False):
        super().__init__()
        assert softmax and to_onehot_y and not sigmoid and not squared_pred and not jaccard and reduction == "mean"
        self.include_background, self.batch = include_background, batch
        self.smooth_nr, self.smooth_dr = float(smooth_nr), float(smooth_dr)

    def forward(self, input, target):
        n_ch = input.shape[1]
        prob = torch.softmax(input, 1)
        idx = target.long()
        onehot = torch.zeros_like(prob).scatter_(1, idx, 1.0)
        if not self.include_background:
            prob, onehot = prob[:, 1:], onehot[:, 1:]
        axes = list(range(2, input.dim()))
        if self.batch:
            axes = [0] + axes
        inter = torch.sum(onehot * prob, dim=axes)
        denom = torch.sum(onehot, dim=axes) + torch.sum(prob, dim=axes)
        f = 1.0 - (2.0 * inter + self.smooth_nr) / (denom + self.smooth_dr)
        return torch.mean(f)


class FocalLoss(nn.Module):
    def __init__(self, gamma=2.0, weight=None, reduction="mean"):
        super().__init__()
        assert weight is None and reduction == "mean"
        self.gamma = gamma

    def forward(self, input, target):
        if input.dim() != target.dim() or target.shape[1] != 1:
            raise ValueError("target must be an index map with one channel")
        if input.dim() > 2:
            i = input.reshape(input.size(0), input.size(1), -1)
            t = target.reshape(target.size(0), 1, -1)
        else:
            i, t = input.unsqueeze(2), target.unsqueeze(2)
        logpt = F.log_softmax(i, dim=1).gather(1, t.long()).squeeze(1)
        pt = torch.exp(logpt)
        w = torch.pow(1.0 - pt, self.gamma)
        return torch.mean(-w * logpt, dim=1).mean()
