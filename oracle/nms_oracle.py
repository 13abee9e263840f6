"""oracle/nms_oracle.py -- CPU restatement of the reference's test-time post-processing.  TEST INFRASTRUCTURE:
imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by the product path.

Restates
  * lib/utils/cython_nms.pyx:37-87   (`nms`: greedy box NMS, "+1" areas, float32, suppress when ovr >= thresh)
  * lib/utils/mask_eval_utils.py:57-110 (`mask_results_with_nms_and_limit_get_index`: per-class candidate
    filter + NMS, then the DETECTIONS_PER_IM limit over all classes)
  * lib/core/test.py:130-133 over lib/modeling/model_builder.py:60-68 (mean over the K refinement heads of
    (cls * iou)[:, 1:])
Pinned: oracle/make_golden.py runs the reference's own Cython `nms` (compiled as it is into oracle/_ref by
oracle/build_ref_nms.py) on seeded detections and stores inputs + keep lists in tests/golden/box_nms.npz; the
restatement reproduces every keep list exactly.  Score ties: numpy's default argsort is not stable, so the
reference's visiting order of equal scores is unspecified; here (and in the CUDA kernel) equal scores are
visited by descending index.
"""
import numpy as np

F32 = np.float32


def nms(dets, thresh):
    """cython_nms.pyx:37-87.  dets [n,5] float32 (x1,y1,x2,y2,score) -> ascending indices that survive."""
    dets = np.asarray(dets, dtype=F32)
    n = dets.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    x1, y1, x2, y2, sc = (dets[:, i] for i in range(5))
    thresh = F32(thresh)
    areas = (x2 - x1 + F32(1)) * (y2 - y1 + F32(1))                       # :45
    order = np.lexsort((np.arange(n), sc))[::-1]                          # :46, ties: descending index
    suppressed = np.zeros(n, dtype=bool)
    for _i in range(n):                                                   # :63-85
        i = order[_i]
        if suppressed[i]:
            continue
        rest = order[_i + 1:]
        rest = rest[~suppressed[rest]]
        if rest.size == 0:
            continue
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(F32(0), xx2 - xx1 + F32(1))
        h = np.maximum(F32(0), yy2 - yy1 + F32(1))
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr >= thresh]] = True
    return np.where(~suppressed)[0]


def class_keep(boxes, scores, score_thresh, nms_thresh):
    """Per-class body of mask_eval_utils.py:63-79: keep [C, n] uint8."""
    boxes, scores = np.asarray(boxes, F32), np.asarray(scores, F32)
    n, c = scores.shape
    keep = np.zeros((c, n), dtype=np.uint8)
    for j in range(c):
        inds = np.where(scores[:, j] > score_thresh)[0]                   # :64
        dets_j = np.hstack((boxes[inds], scores[inds, j][:, None])).astype(F32, copy=False)   # :68
        keep[j, inds[nms(dets_j, nms_thresh)]] = 1                        # :70-77
    return keep


def results_with_nms_and_limit(scores, boxes, score_thresh=1e-5, nms_thresh=0.3, detections_per_im=100):
    """mask_eval_utils.py:57-110 -> (cls_boxes, cls_inds): lists over the score columns of [k,5] float32
    detections and the proposal indices they came from."""
    scores, boxes = np.asarray(scores, F32), np.asarray(boxes, F32)
    keep = class_keep(boxes, scores, score_thresh, nms_thresh)
    c = scores.shape[1]
    cls_inds = [np.where(keep[j])[0] for j in range(c)]
    cls_boxes = [np.hstack((boxes[i], scores[i, j][:, None])).astype(F32) for j, i in enumerate(cls_inds)]
    if detections_per_im > 0:                                             # :82-93
        image_scores = np.hstack([b[:, -1] for b in cls_boxes])
        if len(image_scores) > detections_per_im:
            image_thresh = np.sort(image_scores)[-detections_per_im]
            for j in range(c):
                k = np.where(cls_boxes[j][:, -1] >= image_thresh)[0]
                cls_boxes[j] = cls_boxes[j][k]
                cls_inds[j] = cls_inds[j][k]
    return cls_boxes, cls_inds


def test_scores(ref_cls, ref_iou):
    """lib/core/test.py:130-133: sum over the K heads of (cls * iou)[:, 1:] in head order, then / K (float32)."""
    k = len(ref_cls)
    s = (np.asarray(ref_cls[0], F32) * np.asarray(ref_iou[0], F32))[:, 1:].copy()
    for i in range(1, k):
        s += (np.asarray(ref_cls[i], F32) * np.asarray(ref_iou[i], F32))[:, 1:]
    s /= k
    return s
