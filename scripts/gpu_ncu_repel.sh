#!/bin/bash
# launch list + ncu --set full of the graded-cloud repel iteration (2M points, BoundaryLayerSpacing)
tag=${1:-run}; dt=${2:-f32}
cmd="python scripts/repel_profile.py $dt 4"
timeout 600 $cmd > gpurun_out/plain_repel_$tag.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_repel_$tag.csv $cmd > /dev/null 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'spacing_eval_kernel|repel_tile_kernel|repel_sweep_kernel' -s 3 -c 3 -f -o gpurun_out/prof_repel_$tag $cmd > gpurun_out/ncu_repel_$tag.log 2>&1
echo "ncu_rc=$?"; cat gpurun_out/plain_repel_$tag.log | tail -2
