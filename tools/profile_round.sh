#!/bin/bash
# Round evidence in one GPU call: the ncu launch list of the real step, `ncu --set full` captures of the dominant
# kernels (cfg2 RoIAlign fwd / bwd, overlap; cfg3 window-tile fwd / bwd; cfg5 box NMS; RoIPool tile kernels; the cfg4
# loss block), all as .ncu-rep under gpurun_out/.
# Afterwards, on the build box:  python tools/profile_collect.py r2   (summaries + profiles/traffic.json)
#   gpurun --timeout 1500 -- 'bash tools/profile_round.sh r2'
TAG=${1:-r2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -k "regex:roi_|mask_|score_|cim_|nchw|pairs" -s 150 -c 120 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also "" > /dev/null 2>&1
$NCU --nvtx --metrics gpu__time_duration.sum -k "regex:roi_|mask_|score_|cim_" -s 150 -c 60 --print-nvtx-rename kernel --csv \
    --log-file gpurun_out/${TAG}_launches_nvtx.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also "" > /dev/null 2>&1
for k in roi_align_fwd_tile roi_align_bwd_tile mask_overlap_tc; do
  $NCU --set full --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/${TAG}_full_$k \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --also "" > /dev/null 2>&1
done
for k in roi_align_fwd_win roi_align_bwd_tile; do
  $NCU --set full --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/${TAG}_full_vgg16_$k \
      python tools/roi_bench.py --backbone vgg16 --iters 2 > /dev/null 2>&1
done
# kernels added in the third session of round 2: box NMS over the live list (cfg5), RoIPool tile kernels, clustered loss block
$NCU --set full --import-source on -k regex:cim_box_nms -s 2 -c 1 -f -o gpurun_out/${TAG}_full_infer_box_nms \
    python tools/infer_bench.py --images 8 --iters 2 > /dev/null 2>&1
for k in roi_pool_fwd_tile roi_pool_bwd_tile; do
  $NCU --set full --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/${TAG}_full_$k \
      python tools/roi_bench.py --pool --iters 2 > /dev/null 2>&1
done
$NCU --set full --import-source on -k regex:cim_head_losses -s 4 -c 1 -f -o gpurun_out/${TAG}_full_coco_head_losses \
    python bench.py --workload cfg4_r50_coco_8x2000_q --steps 1 --warmup 3 --no-cpu-baseline --also "" > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -k "regex:roi_|mask_|score_|cim_" -c 60 --csv \
    --log-file gpurun_out/${TAG}_launches_infer.csv python tools/infer_bench.py --images 8 --iters 2 > /dev/null 2>&1
ls -la gpurun_out/${TAG}_*
