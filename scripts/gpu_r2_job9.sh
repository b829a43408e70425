#!/bin/bash
out=gpurun_out; mkdir -p $out; tag=r2j9
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -k "repel or knn_bit_exact" > $out/pytest_$tag.log 2>&1; echo "pytest_rc=$?" )
tail -4 $out/pytest_$tag.log
for dt in f32 f64; do WTP_REPEL_DEBUG=1 timeout 300 python scripts/graded_launches.py $dt 20 2>&1 | tail -8; done | tee $out/graded_debug_$tag.log
timeout 900 python bench.py --no-cpu --no-e2e > $out/bench_$tag.json 2> $out/bench_$tag.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2j9.json').read())
print('value',round(d['value'],1),'repel',round(d['repel']['ms_per_iter'],3))
for k,v in d['extras'].items(): print(' ',k,{a:v[a] for a in v if a not in ('config','roofline','metric','unit','dtype')})
PY
