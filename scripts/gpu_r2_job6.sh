#!/bin/bash
# Round 2, GPU job 6 (1 GPU): the radius guess at the border of the grid (existing cells, clipped ball): parity, leftovers,
# timings; a sweep of the radius margin.
out=gpurun_out; mkdir -p $out; tag=r2j6
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -k "knn or repel_10 or config3 or radius_csr" > $out/pytest_$tag.log 2>&1; echo "pytest_rc=$?" )
tail -4 $out/pytest_$tag.log
run() { timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --repel-iters 8 --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read())
x=d['extras']
print(round(d['value'],1), 'q', round(d['phases_ms']['ms_query'],3), d['tiled_pass_leftovers'], 'repel', round(d['repel']['ms_per_iter'],3), 'f64', round(x['knn_f64_10M']['ms_per_step'],3),
 'cfg3 f32', round(x['repel_config3_graded_2M_f32']['ms_per_iter'],3), x['repel_config3_graded_2M_f32']['leftovers_last_iter'], 'cfg3 f64', round(x['repel_config3_graded_2M_f64']['ms_per_iter'],3), x['repel_config3_graded_2M_f64']['leftovers_last_iter'], 'cfg4', round(x['radius_config4_graded2d_10M_f64']['ms_per_step'],3))"; }
echo "default: $(run)" | tee -a $out/variants_$tag.log
cd whatsthepoint.jl_b200/csrc
for v in "-DTK_RADIUS_SIGMAS=1.7f" "-DTK_RADIUS_SIGMAS=2.5f"; do
  touch knn_tile.cuh; make -j16 EXTRA_NVFLAGS="$v" > /dev/null 2>&1
  echo "variant [$v]: $(cd ../.. && run)" | tee -a ../../$out/variants_$tag.log
done
touch knn_tile.cuh; make -j16 > /dev/null 2>&1
