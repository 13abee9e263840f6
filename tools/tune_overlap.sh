#!/bin/bash
# tuning run of the overlap kernel variants / ablations (results of ablated runs are garbage: timing only)
cd "$(dirname "$0")/.."
for v in ${VARIANTS:-2 4 5 6 7}; do
  echo "== variant $v"; CIM_OVERLAP_VARIANT=$v timeout 120 python tools/bench_overlap.py tiled 2>&1 | tail -1
done
for v in ${ABL_VARIANTS:-5 6}; do
for a in ${ABLS:-3 4 16 19 23}; do
  echo "== variant $v ablation $a"; CIM_OVERLAP_VARIANT=$v CIM_OVERLAP_ABL=$a CIM_OVERLAP_NOCHECK=1 timeout 120 python tools/bench_overlap.py tiled 2>&1 | tail -1
done
done
