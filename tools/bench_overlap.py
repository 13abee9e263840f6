#!/usr/bin/env python
"""Time cim_mask_overlap alone at cfg2 size (8 images x 2000 proposals, 512x512 masks) for the flat and the
tiled (8 x 16 patch) pixel order, check that both give identical maps and counts, and print the fraction of
K-blocks the tensor-core kernel visits.  OVERLAP_DEBUG_FLAGS=2 (this TOOL's switch, handed to cim_set_debug_flags)
selects the loader-warp pipeline variant (tuning aid)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cim_b200 import _lib, mask_ops, synth
_lib.lib().cim_set_debug_flags(int(os.environ.get("OVERLAP_DEBUG_FLAGS", 0)))
n_img, R, S = int(os.environ.get("N_IMG", 8)), 2000, 512
dev = "cuda:0"
masks = [synth.rasterize(synth.proposal_params(R, S, 1234 + b), device=dev) for b in range(n_img)]
layouts = sys.argv[1:] or ["flat", "tiled"]
ref = None
tiles = sum((R + 255) // 256 - (i >> 1) for i in range((R + 127) // 128)) * n_img * (S * S // 128)
for lay in layouts:
    packed = torch.stack([mask_ops.mask_pack(m, layout=lay) for m in masks])
    kbpr = S // 16 if lay == "tiled" else 0
    for _ in range(5):                                # warm-up: clocks, caches
        mask_ops.mask_overlap(packed, algo="tensor", kb_per_row=kbpr)
    out = mask_ops.mask_overlap(packed, return_counts=True, algo="tensor", kb_per_row=kbpr, return_visited=True)
    if ref is None:
        ref = out
    same = lambda a, b: torch.equal(a.view(torch.int16) if a.dtype == torch.float16 else a,
                                    b.view(torch.int16) if b.dtype == torch.float16 else b)
    if not os.environ.get("CIM_OVERLAP_NOCHECK"):
        assert all(same(a, b) for a, b in zip(out[:4], ref[:4])), f"layout {lay} differs"
    best = 1e9
    for rnd in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            mask_ops.mask_overlap(packed, algo="tensor", kb_per_row=kbpr)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    tf = float(R) ** 2 * S * S * n_img / (best * 1e-3) / 1e12
    ex = out[4] * 2.0 * 128 * 256 * 128 / (best * 1e-3) / 1e12
    print(f"layout {lay}: best {best:.3f} ms  visited {out[4] / tiles:.3f} of the K-blocks  "
          f"{ex:.0f} TOP/s executed  {tf:.0f} TOP/s algorithmic-equivalent (upper triangle)")
