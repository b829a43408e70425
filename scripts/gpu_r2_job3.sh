#!/bin/bash
# Round 2, GPU job 3 (1 GPU): radius / k-NN shard tests on the new build, e2e with and without the 3-byte packing, bench line.
out=gpurun_out; mkdir -p $out; tag=r2j3
( timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "radius or shards or knn_bit_exact or real_geometry or full_size" > $out/pytest_$tag.log 2>&1; echo "pytest_rc=$?" )
tail -4 $out/pytest_$tag.log
for p in 0 1; do
  if [ $p = 1 ]; then export WTP_NO_PACK24=1; else unset WTP_NO_PACK24; fi
  WTP_PIPE_DEBUG=1 timeout 300 python scripts/e2e_probe2.py >> $out/e2e_$tag.log 2>&1
  echo "no_pack24=$p" >> $out/e2e_$tag.log
done
unset WTP_NO_PACK24
grep "e2e=\|no_pack" $out/e2e_$tag.log
timeout 900 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench_rc=$?"; python -c "
import json
d=json.loads(open('$out/bench_$tag.json').read())
print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['phases_ms'], d['e2e']['d2h_bytes_per_step'])
print({k:(v['value'],v.get('ms_per_step',v.get('ms_per_iter'))) for k,v in d['extras'].items()})
"
