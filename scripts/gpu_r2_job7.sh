#!/bin/bash
# Round 2, GPU job 7 (1 GPU): the in-kernel retry of overflowing lists: parity, then a sweep of the lane threshold.
out=gpurun_out; mkdir -p $out; tag=r2j7
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -k "knn or repel_10 or config3 or radius_csr" > $out/pytest_$tag.log 2>&1; echo "pytest_rc=$?" )
tail -4 $out/pytest_$tag.log
run() { timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --repel-iters 8 --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read())
x=d['extras']
print(round(d['value'],1), 'q', round(d['phases_ms']['ms_query'],3), d['tiled_pass_leftovers'], 'repel', round(d['repel']['ms_per_iter'],3), 'f64', round(x['knn_f64_10M']['ms_per_step'],3),
 'cfg3 f32', round(x['repel_config3_graded_2M_f32']['ms_per_iter'],3), x['repel_config3_graded_2M_f32']['leftovers_last_iter'], 'cfg3 f64', round(x['repel_config3_graded_2M_f64']['ms_per_iter'],3), x['repel_config3_graded_2M_f64']['leftovers_last_iter'], 'cfg4', round(x['radius_config4_graded2d_10M_f64']['ms_per_step'],3))"; }
echo "default (6): $(run)" | tee -a $out/variants_$tag.log
cd whatsthepoint.jl_b200/csrc
for v in "-DTK_RETRY_LANES=33" "-DTK_RETRY_LANES=2" "-DTK_RETRY_LANES=12"; do
  touch knn_tile.cuh; make -j16 EXTRA_NVFLAGS="$v" > /dev/null 2>&1
  echo "variant [$v]: $(cd ../.. && run)" | tee -a ../../$out/variants_$tag.log
done
touch knn_tile.cuh; make -j16 > /dev/null 2>&1
