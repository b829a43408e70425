#!/bin/bash
# Round 2, GPU job 14 (1 GPU): counting build (caller indices scattered, records gathered by the placement kernel) — equality test,
# bench, per-kernel times of one step from an ncu launch list.
out=gpurun_out; mkdir -p $out; tag=${1:-r2j14}
( timeout 900 python -m pytest tests/test_gpu_index_build.py tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -x > $out/pytest_$tag.log 2>&1; echo "pytest_rc=$?" ); tail -5 $out/pytest_$tag.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench_rc=$?"
python - $tag <<'PY'
import json, sys
f = "bench_%s.json" % sys.argv[1]
d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "phases", {k: round(v, 3) for k, v in d["phases_ms"].items()},
      "e2e", round(d["e2e"]["ms_per_step"], 2), "repel", round(d["repel"]["ms_per_iter"], 3),
      {k: round(v["ms_per_step"] if "ms_per_step" in v else v["ms_per_iter"], 3) for k, v in d["extras"].items()}, d["parity_check"]["ok"])
PY
cmd="python bench.py --steps 1 --warmup 3 --no-cpu --no-repel --no-extras --no-e2e --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_$tag.csv $cmd > $out/ncu_list_$tag.log 2>&1
echo "ncu_list_rc=$?"
python - $tag <<'PY'
import csv, sys
rows = [r for r in csv.reader(open("gpurun_out/launches_%s.csv" % sys.argv[1])) if len(r) > 10]
h = rows[0]; k = h.index("Kernel Name"); v = h.index("Metric Value")
seq = [(r[k][:60], float(r[v].replace(",", ""))) for r in rows[1:]]
# the last step: from the last bbox_partial launch on
last = max(i for i, (n, _) in enumerate(seq) if "bbox_partial" in n)
for n, t in seq[last:]:
    print(f"{t/1000:9.1f} us  {n}")
PY
