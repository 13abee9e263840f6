"""Test-time post-processing on the GPU (cim_test_scores, cim_box_nms through cim_b200.postproc) against the
fixtures of the reference's Cython `nms` and against the oracle.  Keep lists: bit-exact."""
import os

import numpy as np
import pytest
import torch

from cim_b200 import postproc, synth
from oracle import nms_oracle
from conftest import GOLDEN, cim_case_names

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NPZ = np.load(os.path.join(GOLDEN, "box_nms.npz"))
NAMES = cim_case_names(NPZ)


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("name", NAMES)
def test_box_nms_reference_fixtures_bit_exact(name):
    score_thr, nms_thr = NPZ[f"{name}/params"]
    keep = postproc.box_nms(cuda(NPZ[f"{name}/boxes"]), cuda(NPZ[f"{name}/scores"]), score_thr, nms_thr)
    np.testing.assert_array_equal(keep.cpu().numpy(), NPZ[f"{name}/keep"])


@pytest.mark.parametrize("n,c,seed", [(4000, 8, 3), (1237, 20, 4), (65, 3, 5)])
def test_box_nms_large_matches_oracle(n, c, seed):
    """cfg5-sized proposal sets (4000 boxes: multi-chunk path, non-power-of-two padding)."""
    rs = np.random.RandomState(seed)
    boxes = synth.rois_from_params(synth.proposal_params(n, 512, seed)).numpy()[:, 1:]
    scores = np.stack([(rs.permutation(n) + 1.0) / (n + 1) for _ in range(c)], 1).astype(np.float32)
    scores[rs.rand(n, c) < 0.5] = 0
    keep = postproc.box_nms(cuda(boxes), cuda(scores), 1e-5, 0.3).cpu().numpy()
    np.testing.assert_array_equal(keep, nms_oracle.class_keep(boxes, scores, 1e-5, 0.3))


def test_box_nms_idempotent_and_sorted_property():
    """Size-independent properties: kept boxes of a class never overlap >= thresh with a higher-scoring kept
    box; running NMS again on the survivors keeps all of them."""
    n, c = 3000, 4
    rs = np.random.RandomState(9)
    boxes = synth.rois_from_params(synth.proposal_params(n, 512, 9)).numpy()[:, 1:]
    scores = np.stack([(rs.permutation(n) + 1.0) / (n + 1) for _ in range(c)], 1).astype(np.float32)
    keep = postproc.box_nms(cuda(boxes), cuda(scores), 1e-5, 0.3).cpu().numpy()
    for j in range(c):
        idx = np.nonzero(keep[j])[0]
        again = postproc.box_nms(cuda(boxes[idx]), cuda(scores[idx, j:j + 1]), 1e-5, 0.3).cpu().numpy()
        assert again.all()


def test_results_with_nms_and_limit_matches_oracle():
    boxes, scores = NPZ["coco_r500/boxes"], NPZ["coco_r500/scores"]
    s, b, cls_boxes, cls_inds = postproc.results_with_nms_and_limit(cuda(scores), cuda(boxes), 1e-5, 0.3, 100)
    o_boxes, o_inds = nms_oracle.results_with_nms_and_limit(scores, boxes, 1e-5, 0.3, 100)
    assert len(cls_boxes) == scores.shape[1] + 1 and len(cls_boxes[0]) == 0       # shifted by one, :96-105
    for j in range(scores.shape[1]):
        np.testing.assert_array_equal(cls_inds[j + 1], o_inds[j])
        np.testing.assert_array_equal(cls_boxes[j + 1], o_boxes[j])
    want = np.vstack([o_boxes[j] for j in range(scores.shape[1] - 1)])            # :108 leaves the last class out
    np.testing.assert_array_equal(s, want[:, -1])
    np.testing.assert_array_equal(b, want[:, :-1])


def test_test_scores_matches_oracle_bit_exact():
    k, m, c1 = 3, 517, 21
    rs = np.random.RandomState(1)
    scores = rs.rand(2 + 2 * k, m, c1).astype(np.float32)
    got = postproc.test_scores(cuda(scores), k).cpu().numpy()
    want = nms_oracle.test_scores(list(scores[2:2 + k]), list(scores[2 + k:]))
    np.testing.assert_array_equal(got, want)


def test_errors():
    with pytest.raises(RuntimeError):
        postproc.box_nms(torch.zeros(4, 4), torch.zeros(4, 2))                    # CPU tensors: no CPU path
    with pytest.raises(ValueError):
        postproc.box_nms(torch.zeros(4, 5, device=DEV), torch.zeros(4, 2, device=DEV))


def test_box_nms_batched_equals_per_image():
    """cim_box_nms_batched (one CTA per (class, image)) keeps exactly what the per-image call keeps."""
    B, n, c = 3, 700, 12
    g = torch.Generator().manual_seed(9)
    xy = torch.rand(B, n, 2, generator=g) * 400
    wh = torch.rand(B, n, 2, generator=g) * 120 + 4
    boxes = torch.cat([xy, xy + wh], 2).to(DEV)
    scores = torch.rand(B, n, c, generator=g).to(DEV)
    scores[1, :, 3] = 0                                  # a class without candidates
    keep = postproc.box_nms_batched(boxes, scores, 0.05, 0.3)
    assert keep.shape == (B, c, n)
    for b in range(B):
        assert torch.equal(keep[b], postproc.box_nms(boxes[b], scores[b], 0.05, 0.3))
