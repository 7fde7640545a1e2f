"""GPU tests (-m gpu) of the device-side submission post-processing and scorer (xview2_b200/csrc/postprocess.cu):
integer work, bit-exact against the reference's host formulations (scipy.ndimage.label + per-building np.unique vote,
grey dilation, RowPairCalculator)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _blobs(rng, n, h, w, density):
    coarse = rng.random((n, h // 8, w // 8)) < density
    m = np.kron(coarse, np.ones((1, 8, 8), bool))
    noise = rng.random((n, h, w)) < 0.35
    fg = m & ~(noise & (rng.random((n, h, w)) < 0.3))
    cls = rng.integers(1, 5, (n, h, w)).astype(np.uint8)
    return (fg * cls).astype(np.uint8)


def _vote_reference(post):
    """post_process.py:39-43 verbatim."""
    from scipy.ndimage import label
    post = post.copy()
    components, n = label(post > 0)
    for b in range(1, n + 1):
        labels, counts = np.unique(post[components == b], return_counts=True)
        post[components == b] = labels[np.argmax(counts)]
    return post


@pytest.mark.parametrize("shape,density", [((3, 96, 128), 0.3), ((2, 64, 64), 0.7), ((1, 256, 256), 0.15), ((2, 40, 56), 0.0)])
def test_cc_majority_vote_bit_exact(shape, density):
    from xview2_b200 import ops
    rng = np.random.default_rng(5)
    post = _blobs(rng, *shape, density) if density > 0 else np.zeros(shape, np.uint8)
    if density > 0:  # a snake-like component that needs many union steps, and a component touching the tile border
        post[0, 1, :] = 2
        post[0, 1:shape[1] - 1, shape[2] - 2] = 3
        post[0, shape[1] - 2, 1:shape[2] - 1] = 3
    got = ops.cc_majority_vote(torch.from_numpy(post).cuda()).cpu().numpy()
    for i in range(shape[0]):
        assert np.array_equal(got[i], _vote_reference(post[i])), f"tile {i}"
    # components never leak across tiles of the batch (labels are per tile)
    one = ops.cc_majority_vote(torch.from_numpy(post[0]).cuda()).cpu().numpy()
    assert np.array_equal(one, got[0])


def test_cc_majority_vote_full_tile():
    from xview2_b200 import ops
    from xview2_b200.utils.post_process import majority_vote
    rng = np.random.default_rng(9)
    post = _blobs(rng, 2, 1024, 1024, 0.2)
    got = ops.cc_majority_vote(torch.from_numpy(post).cuda()).cpu().numpy()
    for i in range(2):
        assert np.array_equal(got[i], majority_vote(post[i]))  # vectorised host formulation (itself tested against the loop)


@pytest.mark.parametrize("k", [1, 3, 5, 7])
def test_dilate_square_bit_exact(k):
    from scipy.ndimage import grey_dilation

    from xview2_b200 import ops
    rng = np.random.default_rng(k)
    m = _blobs(rng, 2, 72, 88, 0.1)
    got = ops.dilate_square(torch.from_numpy(m).cuda(), k).cpu().numpy()
    for i in range(2):
        assert np.array_equal(got[i], grey_dilation(m[i], size=(k, k)))


def _row_pair(lp, dp, lt, dt):
    """RowPairCalculator.get_row_pair (xview2_metrics.py:77-92) verbatim on arrays."""
    def tp_fn_fp(pred, targ, c):
        return [np.logical_and(pred == c, targ == c).sum(), np.logical_and(pred != c, targ == c).sum(),
                np.logical_and(pred == c, targ != c).sum()]
    lp_b, lt_b, dt_b = ((x > 0).astype(x.dtype) for x in (lp, lt, dt))
    dp = dp * lp_b
    dp, dt = dp[dt_b == 1], dt[dt_b == 1]
    row = tp_fn_fp(lp_b, lt_b, 1)
    for i in range(1, 5):
        row += tp_fn_fp(dp, dt, i)
    return np.array(row, np.int64)


def test_score_counts_and_scorer_cli(tmp_path):
    from PIL import Image

    from xview2_b200 import ops
    from xview2_b200.utils.xview2_metrics import XviewMetrics, scores_from_counters
    rng = np.random.default_rng(2)
    n = 3
    dt = _blobs(rng, n, 1024, 1024, 0.2)
    lt = (dt > 0).astype(np.uint8)
    flip = rng.random((n, 1024, 1024)) < 0.1
    dp = np.where(flip, rng.integers(0, 5, (n, 1024, 1024)), dt).astype(np.uint8)
    lp = ((dp > 0) ^ (rng.random((n, 1024, 1024)) < 0.03)).astype(np.uint8)
    want = sum(_row_pair(lp[i], dp[i], lt[i], dt[i]) for i in range(n))
    got = ops.score_counts(*(torch.from_numpy(a).cuda() for a in (lp, dp, lt, dt))).cpu().numpy()
    assert np.array_equal(got, want)
    os.makedirs(tmp_path / "predictions")
    os.makedirs(tmp_path / "targets")
    for i in range(n):
        Image.fromarray(lp[i]).save(tmp_path / "predictions" / f"test_localization_{i:05d}_prediction.png")
        Image.fromarray(dp[i]).save(tmp_path / "predictions" / f"test_damage_{i:05d}_prediction.png")
        Image.fromarray(lt[i]).save(tmp_path / "targets" / f"test_localization_{i:05d}_target.png")
        Image.fromarray(dt[i]).save(tmp_path / "targets" / f"test_damage_{i:05d}_target.png")
    m = XviewMetrics.compute_score(str(tmp_path / "predictions"), str(tmp_path / "targets"), str(tmp_path / "score.json"))
    assert m.n_tiles == n and m.counters == want.tolist()
    ref = scores_from_counters(want)
    out = json.load(open(tmp_path / "score.json"))
    assert set(out) == {"score", "damage_f1", "localization_f1", "damage_f1_no_damage", "damage_f1_minor_damage",
                        "damage_f1_major_damage", "damage_f1_destroyed"}
    assert all(abs(out[k] - ref[k]) < 1e-12 for k in ref)
    assert abs(out["score"] - (0.3 * out["localization_f1"] + 0.7 * out["damage_f1"])) < 1e-12
