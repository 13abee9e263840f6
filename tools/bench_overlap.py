#!/usr/bin/env python
"""Time cim_mask_overlap alone at cfg2 size (8 images x 2000 proposals, 512x512 masks).
CIM_OVERLAP_VARIANT selects a tile-shape variant of the tensor-core kernel (tuning aid)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cim_b200 import mask_ops, synth
n_img, R = int(os.environ.get("N_IMG", 8)), 2000
dev = "cuda:0"
packed = torch.stack([mask_ops.mask_pack(synth.rasterize(synth.proposal_params(R, 512, 1234 + b), device=dev))
                      for b in range(n_img)])
variants = sys.argv[1:] or ["0"]
best = {v: 1e9 for v in variants}
ref = None
os.environ["CIM_OVERLAP_VARIANT"] = variants[0]
for _ in range(10):                                   # warm-up: clocks, caches
    mask_ops.mask_overlap(packed, algo="tensor")
for rnd in range(3):                                  # interleaved rounds, keep the best of each variant
    for v in variants:
        os.environ["CIM_OVERLAP_VARIANT"] = v
        out = mask_ops.mask_overlap(packed, return_counts=True, algo="tensor")
        if ref is None:
            ref = out
        assert all(torch.equal(a, b) for a, b in zip(out[2:], ref[2:])), f"variant {v} differs"
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            mask_ops.mask_overlap(packed, algo="tensor")
        e1.record()
        torch.cuda.synchronize()
        best[v] = min(best[v], e0.elapsed_time(e1) / 10)
for v in variants:
    tf = float(R) ** 2 * 512 * 512 * n_img / (best[v] * 1e-3) / 1e12
    print(f"variant {v}: best {best[v]:.3f} ms  {tf:.0f} TFLOP/s-equivalent (upper triangle)")
