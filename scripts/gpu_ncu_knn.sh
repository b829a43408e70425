#!/bin/bash
# ncu --set full of the k-NN kernels only (short command: 1 timed step, no repel, no cpu)
tag=${1:-run}
cmd="python bench.py --steps 1 --warmup 3 --no-cpu --no-repel"
timeout 600 $cmd > gpurun_out/plain_$tag.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'knn_tile_kernel|knn_kernel' -s 6 -c 2 -f -o gpurun_out/prof_$tag $cmd > gpurun_out/ncu_full_$tag.log 2>&1
echo "ncu_rc=$?"; tail -3 gpurun_out/ncu_full_$tag.log
