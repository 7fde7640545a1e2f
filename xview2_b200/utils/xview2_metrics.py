"""xView2 scorer with the reference's entry point (/root/reference/utils/xview2_metrics.py): ``XviewMetrics(pred_dir, targ_dir)``
and ``XviewMetrics.compute_score(pred_dir, targ_dir, out_fp)`` over ``*_localization_<id>_prediction.png`` /
``*_damage_<id>_prediction.png`` and the matching ``*_target.png`` files, writing the same JSON keys.

The per-tile TP / FN / FP counting (RowPairCalculator.get_row_pair, xview2_metrics.py:77-92: building masks, damage prediction
masked by the predicted buildings and scored on target-building pixels only) runs on the GPU as ONE pass per batch of tiles
accumulating 15 integers (xv2_score_counts); PNG decoding runs on a thread pool; F1s, their harmonic mean and the 0.3 / 0.7 score
are scalar host arithmetic (xview2_metrics.py:95-137, 243-252).
"""
import json
import os
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np


def f1_from_counts(tp, fp, fn):
    """F1Recorder (xview2_metrics.py:95-137)."""
    precision = 0 if tp == 0 else tp / (tp + fp)
    recall = 0 if tp == 0 else tp / (tp + fn)
    if precision == 0 or recall == 0:
        return 0
    return (2 * precision * recall) / (precision + recall)


def scores_from_counters(c):
    """c: 15 integers lTP lFN lFP, (dTP dFN dFP) x 4 -> the dict compute_score writes (xview2_metrics.py:243-275)."""
    c = [int(v) for v in c]
    lf1 = f1_from_counts(c[0], c[2], c[1])
    df1s = [f1_from_counts(c[3 * k], c[3 * k + 2], c[3 * k + 1]) for k in range(1, 5)]
    df1 = len(df1s) / sum((x + 1e-6) ** -1 for x in df1s)
    return {"score": 0.3 * lf1 + 0.7 * df1, "damage_f1": df1, "localization_f1": lf1, "damage_f1_no_damage": df1s[0],
            "damage_f1_minor_damage": df1s[1], "damage_f1_major_damage": df1s[2], "damage_f1_destroyed": df1s[3]}


class XviewMetrics:
    def __init__(self, pred_dir, targ_dir, batch=16, workers=8):
        import torch

        from .. import ops

        self.pred_dir, self.targ_dir = Path(pred_dir), Path(targ_dir)
        assert self.pred_dir.is_dir(), f"Could not find prediction directory: '{pred_dir}'"
        assert self.targ_dir.is_dir(), f"Could not find target directory: '{targ_dir}'"
        if not torch.cuda.is_available():
            raise RuntimeError("the scorer counts on the GPU (no CPU fallback)")
        quads = []
        for path in sorted(self.targ_dir.glob("*.png")):
            test_hold, loc_dmg, img_id, target = path.name[:-len(".png")].split("_")
            assert loc_dmg in ("localization", "damage") and target == "target", path
            if loc_dmg == "localization":
                quads.append((self.pred_dir / f"{test_hold}_localization_{img_id}_prediction.png",
                              self.pred_dir / f"{test_hold}_damage_{img_id}_prediction.png",
                              self.targ_dir / f"{test_hold}_localization_{img_id}_target.png",
                              self.targ_dir / f"{test_hold}_damage_{img_id}_target.png"))
        self.n_tiles = len(quads)
        counters = torch.zeros(15, dtype=torch.int64, device="cuda")

        def load(path):
            from PIL import Image
            assert path.is_file(), f"file '{path}' does not exist or is not a file"
            img = np.array(Image.open(path))
            assert img.dtype == np.uint8 and img.shape == (1024, 1024), f"{path} must be a 1024x1024 uint8 image"
            return img

        with ThreadPoolExecutor(workers) as pool:
            for i in range(0, len(quads), batch):
                chunk = quads[i:i + batch]
                imgs = list(pool.map(load, [p for q in chunk for p in q]))
                stack = torch.from_numpy(np.stack(imgs).reshape(len(chunk), 4, 1024, 1024)).cuda()
                assert int(stack.max()) <= 4, "values must be ints 0-4"
                ops.score_counts(stack[:, 0], stack[:, 1], stack[:, 2], stack[:, 3], counters)
        self.counters = counters.cpu().tolist()
        self.results = scores_from_counters(self.counters)
        self.lf1, self.df1, self.score = self.results["localization_f1"], self.results["damage_f1"], self.results["score"]
        self.df1s = [self.results[k] for k in ("damage_f1_no_damage", "damage_f1_minor_damage", "damage_f1_major_damage",
                                               "damage_f1_destroyed")]

    def __repr__(self):
        names = ("No damage     (1) ", "Minor damage  (2) ", "Major damage  (3) ", "Destroyed     (4) ")
        s = f"Localization:\n    Buildings | f1: {self.lf1:.4f}\n\nDamage:\n"
        for name, f in zip(names, self.df1s):
            s += f"    {name} | f1: {f:.4f}\n"
        s += f"    Harmonic mean dmgs | f1: {self.df1:.4f}\n\nScore:\n    Score | f1: {self.score:.4f}"
        return s

    @classmethod
    def compute_score(cls, pred_dir, targ_dir, out_fp):
        self = cls(pred_dir, targ_dir)
        with open(out_fp, "w") as f:
            json.dump(self.results, f)
        print(f"Wrote metrics to {out_fp}")
        return self


if __name__ == "__main__":
    import sys

    XviewMetrics.compute_score(*sys.argv[1:4])
