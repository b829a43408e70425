#!/bin/bash
# Round 2, GPU job 37 (1 GPU): the bench line with the config-#5 extras (100 M points generated on the device).
out=gpurun_out; mkdir -p $out
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu > $out/bench_r2j37.json 2> $out/bench_r2j37.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2j37.json").read().strip().splitlines()[-1])
print("value",round(d["value"],1))
for k,v in d["extras"].items():
    print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in ("roofline","config","metric","unit")} if isinstance(v,dict) else v)
PY
