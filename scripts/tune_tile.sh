#!/bin/bash
# Rebuilds knn.cu / repel.cu with tuning macros on the GPU box and times the k-NN bench for each variant.
cd whatsthepoint.jl_b200/csrc
for v in "" "-DTK_APPEND_NOCLOBBER" "-DTK_APPEND_NOCLOBBER -DTK_SWEEP_UNROLL=4" "-DTK_APPEND_NOCLOBBER -DTK_SWEEP_UNROLL=3" "-DTK_SWEEP_UNROLL=1"; do
  touch knn_tile.cuh
  make -j8 EXTRA_NVFLAGS="$v" > /dev/null 2>&1
  r=$(cd ../.. && python bench.py --steps 8 --warmup 3 --no-cpu --repel-iters 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['phases_ms']['ms_query'],3), round(d['repel']['ms_per_iter'],3), d['tiled_pass_leftovers'])")
  echo "variant [$v]: $r"
done
touch knn_tile.cuh; make -j8 > /dev/null 2>&1
