#!/bin/bash
# Rebuilds knn.cu / repel.cu with tuning macros on the GPU box and times the k-NN bench for each variant.
cd whatsthepoint.jl_b200/csrc
for v in "" "-DTK_CAP_F32=1920" "-DTK_CAP_F32=1664" "-DTK_SWEEP_UNROLL=2" "-DTK_SWEEP_UNROLL=8" "-DTK_RADIUS_SIGMAS=2.2f" "-DTK_RADIUS_SIGMAS=2.8f" "-DTK_RADIUS_SIGMAS=2.0f"; do
  touch knn_tile.cuh
  make -j8 EXTRA_NVFLAGS="$v" > /dev/null 2>&1
  r=$(cd ../.. && python bench.py --steps 8 --warmup 3 --no-cpu --no-repel 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['phases_ms']['ms_query'],3), d['tiled_pass_leftovers'])")
  echo "variant [$v]: $r"
done
touch knn_tile.cuh; make -j8 > /dev/null 2>&1
