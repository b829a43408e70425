#!/bin/bash
# Round 2, GPU job 15 (2 GPUs): the windowed counting build on sharded ranks — multi-GPU tests, in-process multi-device check,
# torchrun checks of the row exchange, bench at N = 2 (with parity_check).
out=gpurun_out; mkdir -p $out; tag=r2j15
( timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_index_build.py -m gpu -q > $out/pytest_$tag.log 2>&1; echo "pytest_rc=$?" )
tail -6 $out/pytest_$tag.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 scripts/multigpu_check.py > $out/multigpu_$tag.log 2>&1; echo "multigpu_rc=$?"; tail -4 $out/multigpu_$tag.log
timeout 600 python scripts/multidevice_check.py > $out/multidevice_$tag.log 2>&1; echo "multidevice_rc=$?"; tail -3 $out/multidevice_$tag.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench2_$tag.json 2> $out/bench2_$tag.err
echo "bench2_rc=$?"; tail -3 $out/bench2_$tag.err
python - <<'PY'
import json
for f in ('gpurun_out/bench2_r2j15.json',):
    d=json.loads([l for l in open(f) if l.startswith('{')][0])
    print(f,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'phases',{k: round(v, 3) for k, v in d["phases_ms"].items()},'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],2),d['e2e']['phases_ms'])
    print(' parity',d['parity_check'])
    print(' repel',{k:d['repel'][k] for k in ('value','ms_per_iter','sweep_ms_per_iter','comm_ms_per_iter')})
    for k,v in d['extras'].items(): print(' ',k,{a:v[a] for a in v if a not in ('config','roofline','metric','unit','dtype')})
PY
