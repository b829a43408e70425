#!/bin/bash
# One GPU-box job: parity tests, the bench line, and the ncu evidence for it.
# usage: scripts/gpu_job.sh <tag> [test|bench|ncu ...]
tag=${1:-run}; shift
steps=${@:-test bench ncu}
out=gpurun_out
mkdir -p $out
for s in $steps; do
  case $s in
    test)
      timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest_rc=$?"; tail -3 $out/pytest_$tag.log ;;
    bench)
      timeout 900 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench_rc=$?"; cut -c1-600 $out/bench_$tag.json ;;
    ref)
      timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "ref_rc=$?"; cut -c1-400 $out/bench_ref_$tag.json ;;
    ncu)
      cmd="python bench.py --steps 2 --warmup 3 --no-cpu --repel-iters 3"
      timeout 600 $cmd > $out/plain_$tag.log 2>&1 &&
      timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv $cmd > $out/ncu_list_$tag.log 2>&1
      echo "ncu_list_rc=$?"
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'knn_tile_kernel|repel_tile_kernel' -s 4 -c 3 -f -o $out/prof_$tag $cmd > $out/ncu_full_$tag.log 2>&1
      echo "ncu_full_rc=$?" ;;
    extra)
      timeout 900 python scripts/bench_extra.py > $out/extra_$tag.log 2>&1; echo "extra_rc=$?"; tail -5 $out/extra_$tag.log ;;
  esac
done
