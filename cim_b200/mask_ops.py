"""Pairwise mask-proposal overlap maps on the GPU.

Replaces the reference's offline `mask_iou` / `mask_asymmetric_iou` (lib/utils/mask_utils.py:6-32,
driven by tools/pre/create_cob_iou.py:43-48 and create_cob_asy_iou.py:43-51) and the two pickle
loads + H2D copies per training step that consume them (lib/modeling/model_builder.py:148-156).

    packed = mask_pack(masks_u8)                 # [..., N, H, W] uint8 -> [..., N, words] int32 bits
    iou_map, asy_iou_map = mask_overlap(packed)  # float16 [..., N, N] each, the reference's dtype
asy_iou_map[i, j] = |m_i & m_j| / |m_j| (how much of proposal j lies inside proposal i).
"""
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


@dataclass
class MaskCrops:
    """Compact wire format of N proposal masks: bounding-box crops, bit-packed, 32-pixel aligned in x.
    words [total] int32, meta [N,4] int32 = (wx0, y0, ww, h), off [N] int64 (word offset of each crop).
    Host (numpy / pinned torch) or device tensors."""
    words: torch.Tensor
    meta: torch.Tensor
    off: torch.Tensor
    height: int
    width: int

    @property
    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.words, self.meta, self.off))


def pack_crops_host(masks):
    """numpy/CPU: uint8 / bool masks [N, H, W] -> MaskCrops (CPU tensors).  What a data-loader worker
    would produce from the COB proposals instead of full-image byte masks."""
    m = np.asarray(masks) != 0
    n, h, w = m.shape
    wpad = (w + 31) // 32 * 32
    rows = np.zeros((n, h, wpad), dtype=bool)
    rows[:, :, :w] = m
    bits = np.packbits(rows, axis=2, bitorder="little").view("<u4").reshape(n, h, wpad // 32)   # word k = px 32k..
    meta = np.zeros((n, 4), np.int32)
    off = np.zeros(n, np.int64)
    chunks, total = [], 0
    for i in range(n):
        ys, xs = np.nonzero(m[i].any(1))[0], np.nonzero(m[i].any(0))[0]
        if len(ys) == 0:
            off[i] = total
            continue
        y0, y1, wx0, wx1 = ys[0], ys[-1] + 1, xs[0] // 32, xs[-1] // 32 + 1
        meta[i] = (wx0, y0, wx1 - wx0, y1 - y0)
        off[i] = total
        c = bits[i, y0:y1, wx0:wx1].reshape(-1)
        chunks.append(c)
        total += c.size
    words = np.concatenate(chunks) if chunks else np.zeros(0, np.uint32)
    return MaskCrops(torch.from_numpy(words.view(np.int32).copy()), torch.from_numpy(meta), torch.from_numpy(off),
                     h, w)


def crops_from_packed_host(packed, height, width):
    """CPU: full bit masks [N, H*W/32] (int32, W % 32 == 0) -> MaskCrops, without going through
    byte masks (used by bench.py to derive the wire format from its synthetic packed masks)."""
    if width % 32:
        raise ValueError("crops_from_packed_host needs width % 32 == 0; use pack_crops_host")
    p = np.ascontiguousarray(packed.cpu().numpy() if isinstance(packed, torch.Tensor) else packed)
    p = p.view(np.uint32).reshape(p.shape[0], height, width // 32)
    n = p.shape[0]
    nz = p != 0
    row_any, col_any = nz.any(2), nz.any(1)
    y0 = row_any.argmax(1)
    y1 = height - row_any[:, ::-1].argmax(1)
    x0 = col_any.argmax(1)
    x1 = width // 32 - col_any[:, ::-1].argmax(1)
    empty = ~row_any.any(1)
    meta = np.stack([x0, y0, x1 - x0, y1 - y0], 1).astype(np.int32)
    meta[empty] = 0
    sizes = meta[:, 2].astype(np.int64) * meta[:, 3]
    off = np.zeros(n, np.int64)
    off[1:] = np.cumsum(sizes)[:-1]
    words = np.empty(int(sizes.sum()), np.uint32)
    for i in range(n):
        if sizes[i]:
            words[off[i]:off[i] + sizes[i]] = p[i, y0[i]:y1[i], x0[i]:x1[i]].reshape(-1)
    return MaskCrops(torch.from_numpy(words.view(np.int32)), torch.from_numpy(meta), torch.from_numpy(off),
                     height, width)


def tiled_ok(height, width):
    """The 8 x 16 patch pixel order (cim_mask_pack_tiled) needs whole patches."""
    return height % 8 == 0 and width % 16 == 0


def _tag(packed, kb_per_row):
    packed.cim_kb_per_row = kb_per_row      # read back by mask_overlap (plain attribute, lost on views / copies)
    return packed


def unpack_crops(crops, out=None, layout="auto"):
    """MaskCrops on the device -> full bit masks [N, ceil(H*W/32)] int32 (the input of mask_overlap).
    layout: "flat" (row-major pixels), "tiled" (8 x 16 patches) or "auto" (tiled whenever H % 8 == W % 16 == 0)."""
    _lib.require_cuda(crops.words, "crops.words", torch.int32)
    _lib.require_cuda(crops.meta, "crops.meta", torch.int32)
    _lib.require_cuda(crops.off, "crops.off", torch.int64)
    n = crops.meta.shape[0]
    words = (crops.height * crops.width + 31) // 32
    dev = crops.words.device
    if out is None:
        out = torch.empty((n, words), dtype=torch.int32, device=dev)
    tiled = _want_tiled(layout, crops.height, crops.width)
    L = _lib.lib()
    fn = L.cim_mask_unpack_crops_tiled if tiled else L.cim_mask_unpack_crops
    with torch.cuda.device(dev):
        rc = fn(_lib.ptr(crops.words), _lib.ptr(crops.meta.contiguous()), _lib.ptr(crops.off.contiguous()),
                _lib.ptr(out), n, crops.height, crops.width, words, _lib.stream_ptr(dev))
    _lib.check(rc, "cim_mask_unpack_crops")
    return _tag(out, crops.width // 16 if tiled else 0)


def _want_tiled(layout, height, width):
    if layout not in ("auto", "flat", "tiled"):
        raise ValueError("layout must be 'auto', 'flat' or 'tiled'")
    if layout == "tiled" and not tiled_ok(height, width):
        raise ValueError("the tiled layout needs H % 8 == 0 and W % 16 == 0")
    return layout == "tiled" or (layout == "auto" and tiled_ok(height, width))


def mask_pack(masks, layout="auto"):
    """uint8/bool masks [..., H, W] (non-zero = inside) -> bit masks [..., ceil(H*W/32)] int32.
    layout: "flat" = pixel p = y*W + x is bit p & 31 of word p >> 5; "tiled" = 8 x 16 pixel patches (see
    include/cimhead.h), which lets the tensor-core overlap kernel skip the empty parts of the image; "auto" =
    tiled whenever H % 8 == W % 16 == 0.  The overlap maps do not depend on the layout."""
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
    height, width = masks.shape[-2], masks.shape[-1]
    tiled = _want_tiled(layout, height, width)
    with torch.cuda.device(masks.device):
        if tiled:
            rc = _lib.lib().cim_mask_pack_tiled(_lib.ptr(masks), _lib.ptr(packed), n, height, width, words,
                                                _lib.stream_ptr(masks.device))
        else:
            rc = _lib.lib().cim_mask_pack(_lib.ptr(masks), _lib.ptr(packed), n, hw, words,
                                          _lib.stream_ptr(masks.device))
    _lib.check(rc, "cim_mask_pack")
    return _tag(packed, width // 16 if tiled else 0)


def mask_overlap(packed, return_counts=False, algo="auto", kb_per_row=None, return_visited=False):
    """Bit masks [N, words] or [n_img, N, words] (int32) -> (iou_map, asy_iou_map) float16
    [.., N, N]; with return_counts also (inter int32 [.., N, N], area int32 [.., N]).
    algo: "auto" | "popc" (AND + POPC kernel) | "tensor" (tcgen05 int8 kernel).
    kb_per_row: W // 16 for masks in the tiled pixel order, 0 for flat; None = what mask_pack / unpack_crops
    recorded on the tensor (0 if nothing was).  It only steers the locality sort of the tensor path.
    return_visited: also return the number of K-blocks the tensor path visited (int, 0 on the popc path)."""
    _lib.require_cuda(packed, "packed", torch.int32)
    if kb_per_row is None:
        kb_per_row = int(getattr(packed, "cim_kb_per_row", 0))
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
        ws = torch.empty(L.cim_mask_overlap_workspace_bytes(n_img, n, words, int(return_counts)), dtype=torch.uint8,
                         device=dev)
        if return_visited:
            ws[:8].zero_()
        rc = L.cim_mask_overlap_ex(_lib.ptr(packed), n_img, n, words, int(kb_per_row), _lib.ptr(inter),
                                   _lib.ptr(area), _lib.ptr(iou), _lib.ptr(asy), _lib.ptr(ws), ws.numel(),
                                   _lib.OVERLAP_ALGOS[algo], _lib.stream_ptr(dev))
    _lib.check(rc, "cim_mask_overlap_ex")
    outs = (iou, asy, inter, area) if return_counts else (iou, asy)
    outs = tuple(o.squeeze(0) for o in outs) if squeeze else outs
    if return_visited:
        outs = outs + (int(ws[:8].view(torch.int64).item()),)
    return outs


def mask_overlap_maps(masks):
    """Convenience: uint8 masks [N, H, W] -> (iou_map, asy_iou_map), i.e. what the reference
    un-pickles from cfg.iou_dir / cfg.asy_iou_dir for one image."""
    return mask_overlap(mask_pack(masks))
