"""cim_b200 -- B200 (sm_100a) implementation of the proposal-level hot path of CIM
(ZechengLi19/CIM): RoIAlign / RoIPool fwd+bwd, pairwise mask IoU + containment maps, the scoring
heads, and complete-instance mining / mask NMS / pseudo-label assignment.

Everything here is a thin Python mirror of the reference's operator surface over the C ABI in
include/cimhead.h (libcimhead.so, hand-written CUDA).  There is no CPU fallback.

    from cim_b200.ops import RoIAlign, RoIPool, roi_align, roi_pool      # mmcv.ops surface
    from cim_b200.heads import cls_iou_model, CIM_layer, mine_and_assign  # lib/modeling/heads.py surface
    from cim_b200.mask_ops import mask_pack, mask_overlap                 # lib/utils/mask_utils.py
"""
__version__ = "0.1.0"
