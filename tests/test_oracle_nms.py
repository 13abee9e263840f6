"""The NMS / test-time restatement (oracle/nms_oracle.py) against the fixtures the reference's own Cython
`nms` produced (tests/golden/box_nms.npz, written by oracle/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import nms_oracle
from conftest import GOLDEN, cim_case_names

NPZ = np.load(os.path.join(GOLDEN, "box_nms.npz"))
NAMES = cim_case_names(NPZ)


@pytest.mark.parametrize("name", NAMES)
def test_class_keep_matches_reference(name):
    score_thr, nms_thr = NPZ[f"{name}/params"]
    keep = nms_oracle.class_keep(NPZ[f"{name}/boxes"], NPZ[f"{name}/scores"], score_thr, nms_thr)
    np.testing.assert_array_equal(keep, NPZ[f"{name}/keep"])


def test_threshold_is_inclusive():
    """cython_nms.pyx:84 suppresses when ovr >= thresh: the two fixtures differ by one ulp of the threshold."""
    assert NPZ["exact_third/keep"].sum() < NPZ["above_third/keep"].sum() == NPZ["above_third/keep"].size


def test_limit_keeps_the_top_detections():
    boxes, scores = NPZ["voc_r300/boxes"], NPZ["voc_r300/scores"]
    cls_boxes, cls_inds = nms_oracle.results_with_nms_and_limit(scores, boxes, 1e-5, 0.3, 100)
    all_scores = np.hstack([b[:, -1] for b in cls_boxes])
    assert len(all_scores) == 100                       # tie-free scores: exactly DETECTIONS_PER_IM survive
    full, _ = nms_oracle.results_with_nms_and_limit(scores, boxes, 1e-5, 0.3, 0)
    every = np.sort(np.hstack([b[:, -1] for b in full]))
    np.testing.assert_array_equal(np.sort(all_scores), every[-100:])
    for j, idx in enumerate(cls_inds):
        np.testing.assert_array_equal(cls_boxes[j][:, :4], boxes[idx])


def test_test_scores_is_the_head_mean():
    rs = np.random.RandomState(0)
    cls = [rs.rand(7, 5).astype(np.float32) for _ in range(3)]
    iou = [rs.rand(7, 5).astype(np.float32) for _ in range(3)]
    got = nms_oracle.test_scores(cls, iou)
    want = sum((c.astype(np.float64) * i)[:, 1:] for c, i in zip(cls, iou)) / 3
    np.testing.assert_allclose(got, want, rtol=1e-6)
