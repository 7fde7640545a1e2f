"""Lightning-style ``Model`` with the reference's hooks (/root/reference/model/plt.py:20-179) -- placeholder header;
the full class follows below once the trainer lands.  compute_loss is used by the parity tests today."""


def compute_loss(loss_fn, preds, label, deep_supervision):
    """Model.compute_loss (plt.py:69-77).  The 1, 1/2, 1/4 weights, the 1/(2 - 2^-3) normalisation and the nearest
    label down-sampling are folded into the loss kernels (weight / label stride arguments)."""
    if not deep_supervision:
        return loss_fn(preds, label)
    c_norm = 1 / (2 - 2 ** (-len(preds)))
    loss = loss_fn(preds[0], label, weight=c_norm)
    for i, pred in enumerate(preds[1:]):
        stride = label.shape[-1] // pred.shape[-1]
        loss = loss + loss_fn(pred, label, weight=c_norm * 0.5 ** (i + 1), label_stride=stride)
    return loss
