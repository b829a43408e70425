#!/bin/bash
# Round 2, GPU job 4 (8 GPUs): the bench line at N = 8 (row exchange over 8 peers, parity_check on every rank), then the
# single-process context over 4 devices.
out=gpurun_out; mkdir -p $out; tag=r2j11
nvidia-smi -L > $out/host_$tag.txt; nproc >> $out/host_$tag.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 8 --steps 10 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err
echo "bench_rc=$?"; tail -3 $out/bench_$tag.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2j11.json') if l.startswith('{')][0])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['phases_ms'])
print('parity',d['parity_check'])
print('repel',{k:d['repel'][k] for k in ('value','ms_per_iter','sweep_ms_per_iter','comm_ms_per_iter')})
print({k:(v['value'],v.get('ms_per_step',v.get('ms_per_iter'))) for k,v in d['extras'].items()})
print(d['phases_ms'], d['index_window'])
PY
timeout 600 python scripts/multidevice_check.py 4 > $out/multidevice4_$tag.log 2>&1; echo "multidevice_rc=$?"; grep -c PASS $out/multidevice4_$tag.log; grep "FAIL\|wtp_knn_f32 10M\|MULTIDEVICE" $out/multidevice4_$tag.log
