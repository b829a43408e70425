#!/bin/bash
# Round 2, GPU job 5 (2 GPUs): density classes of the graded repel, sliced upload + staged row exchange, bench at N = 2 and N = 1.
out=gpurun_out; mkdir -p $out; tag=r2j5
( timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -k "multi or sharded or repel or shards" > $out/pytest_$tag.log 2>&1; echo "pytest_rc=$?" )
tail -6 $out/pytest_$tag.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench2_$tag.json 2> $out/bench2_$tag.err
echo "bench2_rc=$?"; tail -3 $out/bench2_$tag.err
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py > $out/bench1_$tag.json 2> $out/bench1_$tag.err; echo "bench1_rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/bench2_r2j5.json','gpurun_out/bench1_r2j5.json'):
    d=json.loads([l for l in open(f) if l.startswith('{')][0])
    print(f,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],2),d['e2e']['phases_ms'], d['e2e']['d2h_bytes_per_step'])
    print(' parity',d['parity_check'])
    print(' repel',{k:d['repel'][k] for k in ('value','ms_per_iter','sweep_ms_per_iter','comm_ms_per_iter')})
    for k,v in d['extras'].items(): print(' ',k,{a:v[a] for a in v if a not in ('config','roofline','metric','unit','dtype')})
PY
