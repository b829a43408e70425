#!/bin/bash
# Round 2, GPU job 24 (1 GPU): final evidence on the final build — whole -m gpu suite, the bench line, launch list, ncu --set full of
# the two tiled kernels, DRAM traffic per step of the extra configurations.
out=gpurun_out; mkdir -p $out; tag=r2j24
( timeout 2400 python -m pytest tests -m gpu -q > $out/pytest_$tag.log 2>&1; echo "pytest_rc=$?" ); tail -4 $out/pytest_$tag.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench_rc=$?"; cut -c1-250 $out/bench_$tag.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "ref_rc=$?"; cut -c1-300 $out/bench_ref_$tag.json
cmd="python bench.py --steps 2 --warmup 3 --no-cpu --repel-iters 3 --no-extras --no-e2e"
timeout 600 $cmd > $out/plain_$tag.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv $cmd > $out/ncu_list_$tag.log 2>&1
echo "ncu_list_rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn_tile_kernel|repel_tile_kernel' -s 4 -c 3 -f -o $out/prof_$tag $cmd > $out/ncu_full_$tag.log 2>&1
echo "ncu_full_rc=$?"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for w in "knn_f64 3" "cfg3_f32 8" "cfg3_f64 8" "cfg4 3"; do set -- $w
  timeout 300 python scripts/ncu_extras.py $1 $2 > /dev/null 2>&1 &&
  timeout 900 ncu --metrics $M --clock-control none --csv --log-file $out/extra_${1}_$tag.csv python scripts/ncu_extras.py $1 $2 > /dev/null 2>&1
  echo "extra $1 rc=$?"
done
