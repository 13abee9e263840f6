"""Pairwise mask-proposal overlap maps on the GPU.

Replaces the reference's offline `mask_iou` / `mask_asymmetric_iou` (lib/utils/mask_utils.py:6-32,
driven by tools/pre/create_cob_iou.py:43-48 and create_cob_asy_iou.py:43-51) and the two pickle
loads + H2D copies per training step that consume them (lib/modeling/model_builder.py:148-156).

    packed = mask_pack(masks_u8)                 # [..., N, H, W] uint8 -> [..., N, words] int32 bits
    iou_map, asy_iou_map = mask_overlap(packed)  # float16 [..., N, N] each, the reference's dtype
asy_iou_map[i, j] = |m_i & m_j| / |m_j| (how much of proposal j lies inside proposal i).
"""
import torch

from . import _lib


def mask_pack(masks):
    """uint8/bool masks [..., H, W] (non-zero = inside) -> bit masks [..., ceil(H*W/32)] int32."""
    _lib.require_cuda(masks, "masks")
    if masks.dtype == torch.bool:
        masks = masks.view(torch.uint8)
    if masks.dtype != torch.uint8:
        raise TypeError("masks must be uint8 or bool")
    if masks.dim() < 3:
        raise ValueError("masks must be [..., H, W]")
    masks = masks.contiguous()
    lead = masks.shape[:-2]
    hw = masks.shape[-1] * masks.shape[-2]
    n = 1
    for s in lead:
        n *= s
    words = (hw + 31) // 32
    packed = torch.empty(lead + (words,), dtype=torch.int32, device=masks.device)
    with torch.cuda.device(masks.device):
        rc = _lib.lib().cim_mask_pack(_lib.ptr(masks), _lib.ptr(packed), n, hw, words,
                                      _lib.stream_ptr(masks.device))
    _lib.check(rc, "cim_mask_pack")
    return packed


def mask_overlap(packed, return_counts=False, algo="auto"):
    """Bit masks [N, words] or [n_img, N, words] (int32) -> (iou_map, asy_iou_map) float16
    [.., N, N]; with return_counts also (inter int32 [.., N, N], area int32 [.., N]).
    algo: "auto" | "popc" (AND + POPC kernel) | "tensor" (tcgen05 int8 kernel)."""
    _lib.require_cuda(packed, "packed", torch.int32)
    squeeze = packed.dim() == 2
    if squeeze:
        packed = packed.unsqueeze(0)
    if packed.dim() != 3:
        raise ValueError("packed must be [N, words] or [n_img, N, words]")
    packed = packed.contiguous()
    n_img, n, words = packed.shape
    dev = packed.device
    L = _lib.lib()
    iou = torch.empty((n_img, n, n), dtype=torch.float16, device=dev)
    asy = torch.empty((n_img, n, n), dtype=torch.float16, device=dev)
    area = torch.empty((n_img, n), dtype=torch.int32, device=dev)
    inter = torch.empty((n_img, n, n), dtype=torch.int32, device=dev) if return_counts else None
    with torch.cuda.device(dev):
        rc = L.cim_mask_overlap_algo(_lib.ptr(packed), n_img, n, words, _lib.ptr(inter), _lib.ptr(area),
                                     _lib.ptr(iou), _lib.ptr(asy), None, 0, _lib.OVERLAP_ALGOS[algo],
                                     _lib.stream_ptr(dev))
    _lib.check(rc, "cim_mask_overlap_algo")
    outs = (iou, asy, inter, area) if return_counts else (iou, asy)
    return tuple(o.squeeze(0) for o in outs) if squeeze else outs


def mask_overlap_maps(masks):
    """Convenience: uint8 masks [N, H, W] -> (iou_map, asy_iou_map), i.e. what the reference
    un-pickles from cfg.iou_dir / cfg.asy_iou_dir for one image."""
    return mask_overlap(mask_pack(masks))
