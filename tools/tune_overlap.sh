#!/bin/bash
# tuning run of the overlap kernel variants / ablations (results of ablated runs are garbage: timing only)
cd "$(dirname "$0")/.."
for v in 2 4 3; do
  echo "== variant $v"; CIM_OVERLAP_VARIANT=$v timeout 120 python tools/bench_overlap.py tiled 2>&1 | tail -1
done
for a in 1 2 3 4 16 19; do
  echo "== variant 3 ablation $a"; CIM_OVERLAP_VARIANT=3 CIM_OVERLAP_ABL=$a CIM_OVERLAP_NOCHECK=1 timeout 120 python tools/bench_overlap.py tiled 2>&1 | tail -1
done
