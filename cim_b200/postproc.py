"""Test-time post-processing of the CIM heads on the GPU, with the reference's surface:

  test_scores(scores, k)                       lib/core/test.py:130-133 (mean over the refinement heads of
                                               model_builder.testing_function's (cls * iou)[:, 1:])
  box_nms(boxes, scores, score_thresh, nms)    per-class candidate filter + lib/utils/cython_nms.pyx `nms`
  results_with_nms_and_limit(...)              lib/utils/mask_eval_utils.py:57-110
                                               (mask_results_with_nms_and_limit_get_index)

All arithmetic runs in libcimhead.so (cim_test_scores / cim_box_nms); there is no CPU path.
"""
import numpy as np
import torch

from . import _lib


def test_scores(scores, k):
    """scores [2+2K, M, C+1] (cls_iou_model.forward_batched) -> [M, C]: mean_k (ref_cls_k * ref_iou_k)[:, 1:]."""
    _lib.require_cuda(scores, "scores", torch.float32)
    scores = scores.contiguous()
    nh, m, c1 = scores.shape
    if nh != 2 + 2 * k:
        raise ValueError("scores must hold 2 + 2K heads")
    with torch.cuda.device(scores.device):
        out = torch.empty((m, c1 - 1), dtype=torch.float32, device=scores.device)
        rc = _lib.lib().cim_test_scores(_lib.ptr(scores), _lib.ptr(out), m, c1, k, _lib.stream_ptr(scores.device))
    _lib.check(rc, "cim_test_scores")
    return out


def box_nms(boxes, scores, score_thresh=1e-5, nms_thresh=0.3):
    """boxes [n,4] (x1,y1,x2,y2), scores [n,C] -> keep [C,n] uint8: proposal i survives class c's NMS."""
    _lib.require_cuda(boxes, "boxes", torch.float32)
    _lib.require_cuda(scores, "scores", torch.float32)
    boxes, scores = boxes.contiguous(), scores.contiguous()
    n, c = scores.shape
    if boxes.shape != (n, 4):
        raise ValueError("boxes must be [n, 4]")
    with torch.cuda.device(boxes.device):
        keep = torch.empty((c, n), dtype=torch.uint8, device=boxes.device)
        rc = _lib.lib().cim_box_nms(_lib.ptr(boxes), _lib.ptr(scores), n, c, c, float(score_thresh),
                                    float(nms_thresh), _lib.ptr(keep), _lib.stream_ptr(boxes.device))
    _lib.check(rc, "cim_box_nms")
    return keep


def box_nms_batched(boxes, scores, score_thresh=1e-5, nms_thresh=0.3):
    """boxes [B,n,4], scores [B,n,C] -> keep [B,C,n] uint8: box_nms for every image of a batch in one launch."""
    _lib.require_cuda(boxes, "boxes", torch.float32)
    _lib.require_cuda(scores, "scores", torch.float32)
    boxes, scores = boxes.contiguous(), scores.contiguous()
    b, n, c = scores.shape
    if boxes.shape != (b, n, 4):
        raise ValueError("boxes must be [B, n, 4]")
    with torch.cuda.device(boxes.device):
        keep = torch.empty((b, c, n), dtype=torch.uint8, device=boxes.device)
        rc = _lib.lib().cim_box_nms_batched(_lib.ptr(boxes), _lib.ptr(scores), b, n, c, c, float(score_thresh),
                                            float(nms_thresh), _lib.ptr(keep), _lib.stream_ptr(boxes.device))
    _lib.check(rc, "cim_box_nms_batched")
    return keep


def results_with_nms_and_limit(scores, boxes, score_thresh=1e-5, nms_thresh=0.3, detections_per_im=100):
    """mask_results_with_nms_and_limit_get_index (mask_eval_utils.py:57-110) for one image.
    scores [n,C], boxes [n,4] CUDA tensors.  Returns (scores, boxes, cls_boxes, cls_inds) with the reference's
    conventions: cls_boxes / cls_inds are lists of length C + 1 shifted by one (entry 0 empty, :96-105), holding
    numpy [k,5] float32 detections and the proposal indices they came from; the flat `scores` / `boxes` stack
    entries 1 .. C-1 exactly as :108-110 does (the last class is left out there too)."""
    keep = box_nms(boxes, scores, score_thresh, nms_thresh)
    c = scores.shape[1]
    kept_scores = torch.where(keep.t().bool(), scores, torch.full_like(scores, float("-inf")))
    total = int(keep.sum().item())
    if detections_per_im > 0 and total > detections_per_im:               # :82-93
        thresh = torch.topk(kept_scores.flatten(), detections_per_im).values[-1]
        keep = keep & (kept_scores.t() >= thresh).to(torch.uint8)
    keep_h, scores_h, boxes_h = keep.cpu().numpy(), scores.cpu().numpy(), boxes.cpu().numpy()
    cls_boxes, cls_inds = [[]], [[]]
    for j in range(c):
        idx = np.nonzero(keep_h[j])[0]
        cls_inds.append(idx)
        cls_boxes.append(np.hstack((boxes_h[idx], scores_h[idx, j][:, None])).astype(np.float32, copy=False))
    im_results = np.vstack([cls_boxes[j] for j in range(1, c)]) if c > 1 else np.zeros((0, 5), np.float32)
    return im_results[:, -1], im_results[:, :-1], cls_boxes, cls_inds
