#!/bin/bash
# Rebuilds repel.cu with tuning macros on the GPU box and times the repel part of the bench for each variant.
cd whatsthepoint.jl_b200/csrc
for v in "" "-DWTP_NO_INLINE_CLIPPED"; do
  touch repel.cu
  make -j8 EXTRA_NVFLAGS="$v" > /dev/null 2>&1
  r=$(cd ../.. && python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --repel-iters 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['repel']['ms_per_iter'],3), round(d['repel']['sweep_ms_per_iter'],3), d['repel']['conv_last'])")
  echo "variant [$v]: $r"
done
touch repel.cu; make -j8 > /dev/null 2>&1
