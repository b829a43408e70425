#!/bin/bash
# Round 2, GPU job 25 (8 GPUs): the bench lines of the final build at N = 8 and N = 4 (what the driver's scaling run launches),
# with parity_check on every rank.
out=gpurun_out; mkdir -p $out; tag=r2j25
nvidia-smi -L > $out/host_$tag.txt; nproc >> $out/host_$tag.txt
for n in ${NLIST:-8}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 10 --warmup 3 > $out/bench${n}_$tag.json 2> $out/bench${n}_$tag.err
  echo "bench${n}_rc=$?"; tail -2 $out/bench${n}_$tag.err
done
python - <<'PY'
import json
for n in (8, 4)[:1]:
    try: d=json.loads([l for l in open('gpurun_out/bench%d_r2j25.json' % n) if l.startswith('{')][0])
    except Exception as e: print(n, 'unreadable', e); continue
    print(n, 'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],2),d['e2e']['phases_ms'])
    print(' parity',d['parity_check'])
    print(' repel',{k:d['repel'][k] for k in ('value','ms_per_iter','sweep_ms_per_iter','comm_ms_per_iter')})
    print(' ',{k:(round(v['value'],1),round(v.get('ms_per_step',v.get('ms_per_iter')),3)) for k,v in d['extras'].items()})
    print(' ',d['phases_ms'], d['index_window'])
PY
