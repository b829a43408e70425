#!/bin/bash
# Round 2, GPU job 31 (2 GPUs): where the sharded e2e step time goes — per-rank step times of the timed loop and of one profiled call.
out=gpurun_out; mkdir -p $out; tag=r2j31
nproc > $out/host_$tag.txt; numactl -H >> $out/host_$tag.txt 2>&1; nvidia-smi topo -m >> $out/host_$tag.txt 2>&1
for v in "" "WTP_STAGE_SLOTS=16"; do
  env $v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 10 --warmup 3 --no-repel --no-extras --no-cpu > $out/bench2_$tag.json 2> $out/bench2_$tag.err
  echo "rc=$? [$v]"
  python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench2_r2j31.json') if l.startswith('{')][0])
print('value',round(d['value'],1),'e2e',round(d['e2e']['ms_per_step'],2),d['e2e']['ms_per_step_per_rank'],d['e2e']['phases_ms'])
PY
done
