#!/usr/bin/env python
"""Times the LITERAL reference mask_utils.mask_iou / mask_asymmetric_iou double loop
(/root/reference/lib/utils/mask_utils.py:6-32, driven column by column as tools/pre/create_cob_iou.py:43-48 and
create_cob_asy_iou.py:43-51 do) behind the numpy shim of oracle/make_golden.py, on a small set of synthetic proposal
masks, and writes microseconds per (pair, map) to profiles/literal_mask_utils.json.

/root/reference exists only in the build container, so bench.py cannot run this on the GPU box; it reads the committed
measurement and extrapolates it by the pair count (cpu_baseline.literal_mask_utils_s_per_image).  BASELINE.md section 4
promised this figure next to the vectorised restatement the CPU arm times.

    python tools/literal_mask_utils_timing.py [--n 48] [--mask 512]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=48)
    ap.add_argument("--mask", type=int, nargs="*", default=[512, 128])
    args = ap.parse_args()
    import make_golden
    from cim_b200 import synth
    _, mu = make_golden.load_reference()
    out = {"source": "/root/reference/lib/utils/mask_utils.py (unmodified) behind chainer.backends.cuda.get_array_module -> numpy",
           "driver": "column by column, tools/pre/create_cob_iou.py:43-48 + create_cob_asy_iou.py:43-51",
           "host_cores": os.cpu_count(), "threads_used": 1, "n_masks": args.n, "sizes": {}}
    for side in args.mask:
        params = synth.proposal_params(args.n, 512, 1234)
        masks = synth.rasterize(params, out_size=side).numpy().astype(np.uint8)
        n = len(masks)
        t0 = time.perf_counter()
        iou = np.zeros((n, n), np.float32)
        asy = np.zeros((n, n), np.float32)
        for j in range(n):
            iou[:, j] = mu.mask_iou(masks, masks[j:j + 1])[:, 0]
            asy[:, j] = mu.mask_asymmetric_iou(masks, masks[j:j + 1])[:, 0]
        iou.astype(np.float16), asy.astype(np.float16)
        dt = time.perf_counter() - t0
        out["sizes"][str(side)] = {"seconds": dt, "us_per_pair_both_maps": 1e6 * dt / (n * n)}
        print(f"{side}x{side}: {n} masks, {dt:.2f} s, {1e6 * dt / (n * n):.1f} us per pair (iou + asy)")
    with open(os.path.join(ROOT, "profiles", "literal_mask_utils.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
