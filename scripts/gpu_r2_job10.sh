#!/bin/bash
out=gpurun_out; mkdir -p $out; tag=r2j10
( timeout 2000 python -m pytest tests -m gpu -q > $out/pytest_$tag.log 2>&1; echo "pytest_rc=$?" )
tail -6 $out/pytest_$tag.log
for dt in f32 f64; do WTP_REPEL_DEBUG=1 timeout 300 python scripts/graded_launches.py $dt 20 2>&1 | grep "it 2\|ms_total" | cut -c1-200; done | tee $out/graded_debug_$tag.log
timeout 900 python bench.py --no-cpu > $out/bench_$tag.json 2> $out/bench_$tag.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2j10.json').read())
print('value',round(d['value'],1),d['phases_ms'],d['tiled_pass_leftovers'],'e2e',round(d['e2e']['value'],1),'repel',round(d['repel']['ms_per_iter'],3))
for k,v in d['extras'].items(): print(' ',k,{a:v[a] for a in v if a not in ('config','roofline','metric','unit','dtype')})
PY
