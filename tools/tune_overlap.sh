#!/bin/bash
# A/B timing of the two tensor-core overlap kernels at cfg2 size: 2 = loader warp + staging ring, 0 = default (direct)
cd "$(dirname "$0")/.."
for v in ${VARIANTS:-2 0}; do
  echo "== debug flags $v"; OVERLAP_DEBUG_FLAGS=$v timeout 120 python tools/bench_overlap.py tiled 2>&1 | tail -1
done
