#!/bin/bash
# Round 2, GPU job 1: parity (all -m gpu tests), the bench line, A/B of the filter loop width and the staging slot
# size, launch list + one ncu --set full capture of the tiled kernels.
out=gpurun_out; mkdir -p $out; tag=r2j1
nproc > $out/host_$tag.txt; lscpu | egrep 'Model name|Socket|Thread|^CPU\(s\)|L3' >> $out/host_$tag.txt; nvidia-smi -L >> $out/host_$tag.txt
( timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest_rc=$?" ) 
tail -5 $out/pytest_$tag.log
timeout 900 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench_rc=$?"; cut -c1-400 $out/bench_$tag.json; tail -3 $out/bench_$tag.err
# e2e: staging slot size / ring depth
for cfg in "32 3" "8 4" "4 8" "2 16" "1 32"; do set -- $cfg
  WTP_STAGE_MB=$1 WTP_STAGE_SLOTS=$2 WTP_PIPE_DEBUG=1 timeout 300 python scripts/e2e_probe2.py >> $out/e2e_$tag.log 2>&1
done
grep "e2e=" $out/e2e_$tag.log
# filter loop width A/B (rebuild knn.cu / repel.cu on the box)
cd whatsthepoint.jl_b200/csrc
for v in "-DTK_FILTER_WIDE=1" "-DTK_FILTER_WIDE=2" "-DTK_FILTER_WIDE=4"; do
  touch knn_tile.cuh
  make -j16 EXTRA_NVFLAGS="$v" > /dev/null 2>&1
  r=$(cd ../.. && timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --repel-iters 8 --no-extras --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['phases_ms']['ms_query'],3), round(d['repel']['ms_per_iter'],3), d['tiled_pass_leftovers'])")
  echo "variant [$v]: $r" | tee -a ../../$out/variants_$tag.log
done
touch knn_tile.cuh; make -j16 > /dev/null 2>&1
cd ../..
cmd="python bench.py --steps 2 --warmup 3 --no-cpu --repel-iters 3 --no-extras --no-e2e"
timeout 600 $cmd > $out/plain_$tag.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv $cmd > $out/ncu_list_$tag.log 2>&1
echo "ncu_list_rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn_tile_kernel|repel_tile_kernel' -s 4 -c 3 -f -o $out/prof_$tag $cmd > $out/ncu_full_$tag.log 2>&1
echo "ncu_full_rc=$?"
