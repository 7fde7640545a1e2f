"""Probabilities -> xView2 submission label maps, the reference's ``utils/post_process.py`` (/root/reference/utils/post_process.py:27-67).

    python -m xview2_b200.utils.post_process --results /results [--components] [--dilate] [--dilation_rate 3]

reads ``<results>/probs/*localization*.npy`` (h, w) and ``*damage*.npy`` (4, h, w) as ``Model.save`` wrote them and writes
``<results>/predictions/*_prediction.png``.

  * the per-pixel rule (``post = argmax(dmg) + 1``; ``pre = loc > 0.3 | (loc > 0.1 & post > 1)``; ``post *= pre``, lines :31-38)
    runs on the GPU (xv2_post_process_probs), batched over files with the .npy reads / PNG writes on a thread pool;
  * ``--components`` (:39-43): every connected building (4-connectivity, scipy.ndimage.label's default like the reference) takes
    its majority damage class; on the device by union-find label equivalence + one (component, class) histogram
    (xv2_cc_majority_vote); ties go to the smallest class like ``np.unique`` + ``argmax``;
  * ``--dilate`` (:44-45): grey-scale dilation with a square footprint (skimage ``dilation(img, square(k))``) on the device
    (xv2_dilate_square, odd k; even k falls back to scipy.ndimage.grey_dilation, the same operator).
``majority_vote`` / ``dilate`` below are the host formulations (numpy / scipy) the device kernels are tested against.
"""
import os
from argparse import ArgumentDefaultsHelpFormatter, ArgumentParser
from concurrent.futures import ThreadPoolExecutor
from glob import glob

import numpy as np


def majority_vote(post):
    """post: (h, w) integer map, 0 = background.  Returns a copy in which every 4-connected component of ``post > 0`` carries
    its most frequent value (post_process.py:39-43)."""
    from scipy.ndimage import label

    components, n = label(post > 0)
    if n == 0:
        return post.copy()
    mask = components > 0
    comp = components[mask].astype(np.int64)
    cls = post[mask].astype(np.int64)
    width = int(cls.max()) + 1
    votes = np.bincount(comp * width + cls, minlength=(n + 1) * width).reshape(n + 1, width)
    winner = votes.argmax(1)  # first maximum = smallest class on ties, like np.unique(...)[np.argmax(counts)]
    out = post.copy()
    out[mask] = winner[comp].astype(post.dtype)
    return out


def dilate(img, size):
    from scipy.ndimage import grey_dilation

    return grey_dilation(img, size=(size, size))


def rule_maps(loc, dmg):
    """(pre, post) uint8 maps from loc (h, w) and dmg (4, h, w) probabilities on the GPU (post_process.py:31-38)."""
    import torch

    from .. import ops

    if not torch.cuda.is_available():
        raise RuntimeError("post-process label maps are computed on the GPU (no CPU fallback)")
    pre, post = ops.post_process_probs(torch.from_numpy(np.ascontiguousarray(loc, np.float32)).cuda(),
                                       torch.from_numpy(np.ascontiguousarray(dmg, np.float32)).cuda())
    return pre.cpu().numpy(), post.cpu().numpy()


def post_process(args, pre_path, post_path, out_dir):
    from PIL import Image

    loc, dmg = np.load(pre_path), np.load(post_path)
    if dmg.ndim == 3 and dmg.shape[0] == 4:
        pre, post = rule_maps(loc, dmg)
    else:  # already a label map (mse / coral heads of the reference): host arithmetic on integers
        post = dmg.astype(np.uint8)
        pre = ((loc > 0.3) | ((loc > 0.1) & (post > 1))).astype(np.uint8)
        post = post * pre
    if args.components or args.dilate:
        import torch

        from .. import ops
        pre_d, post_d = torch.from_numpy(np.ascontiguousarray(pre, np.uint8)).cuda(), torch.from_numpy(np.ascontiguousarray(post, np.uint8)).cuda()
        if args.components:
            post_d = ops.cc_majority_vote(post_d)
        if args.dilate and args.dilation_rate % 2 == 1:
            pre_d, post_d = ops.dilate_square(pre_d, args.dilation_rate), ops.dilate_square(post_d, args.dilation_rate)
        pre, post = pre_d.cpu().numpy(), post_d.cpu().numpy()
        if args.dilate and args.dilation_rate % 2 == 0:
            pre, post = dilate(pre, args.dilation_rate), dilate(post, args.dilation_rate)
    for arr, path in ((pre, pre_path), (post, post_path)):
        Image.fromarray(arr.astype(np.uint8)).save(os.path.join(out_dir, os.path.basename(path).replace(".npy", "_prediction.png")))


def build_parser():
    parser = ArgumentParser(formatter_class=ArgumentDefaultsHelpFormatter)
    parser.add_argument("--results", type=str, default="/results", help="directory holding probs/ (input) and predictions/ (output)")
    parser.add_argument("--components", action="store_true", help="majority damage class per connected building")
    parser.add_argument("--dilate", action="store_true", help="dilate the localisation and damage maps")
    parser.add_argument("--dilation_rate", type=int, default=3, help="side of the square dilation footprint")
    parser.add_argument("--workers", type=int, default=8, help="threads for .npy reads, host steps and PNG writes")
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    out_dir = os.path.join(args.results, "predictions")
    os.makedirs(out_dir, exist_ok=True)
    pre_pred = sorted(glob(os.path.join(args.results, "probs", "*localization*")))
    post_pred = sorted(glob(os.path.join(args.results, "probs", "*damage*")))
    assert len(pre_pred) == len(post_pred), "every localisation probability map needs its damage map"
    with ThreadPoolExecutor(args.workers) as pool:
        list(pool.map(lambda pq: post_process(args, pq[0], pq[1], out_dir), zip(pre_pred, post_pred)))
    return len(pre_pred)


if __name__ == "__main__":
    main()
