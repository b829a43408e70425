#!/bin/bash
# Round 2, GPU job 17 (1 GPU): ncu --set full of the two kernels that bound the extra configurations:
# radius_tile_fill_kernel<double,2> on config #4 and spacing_eval_ordered_kernel<float,3> on config #3.
out=gpurun_out; mkdir -p $out; tag=r2j17
timeout 300 python scripts/ncu_extras.py cfg4 2 > /dev/null 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'radius_tile_fill_kernel|radius_tile_count_kernel' -s 2 -c 2 -f -o $out/prof_cfg4_$tag python scripts/ncu_extras.py cfg4 2 > $out/ncu_cfg4_$tag.log 2>&1
echo "cfg4_rc=$?"
timeout 300 python scripts/ncu_extras.py cfg3_f32 4 > /dev/null 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'spacing_eval_ordered_kernel' -s 2 -c 1 -f -o $out/prof_cfg3_$tag python scripts/ncu_extras.py cfg3_f32 4 > $out/ncu_cfg3_$tag.log 2>&1
echo "cfg3_rc=$?"
