#!/bin/bash
# Round 2, GPU job 13 (1 GPU): counting-sort index build — equality with the radix build, parity suites, bench with either build.
out=gpurun_out; mkdir -p $out; tag=r2j13
( timeout 900 python -m pytest tests/test_gpu_index_build.py tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -x > $out/pytest_$tag.log 2>&1; echo "pytest_rc=$?" ); tail -15 $out/pytest_$tag.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench_rc=$?"
WTP_RADIX_BUILD=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > $out/bench_radix_$tag.json 2> $out/bench_radix_$tag.err; echo "bench_radix_rc=$?"
python - <<'PY'
import json
for f in ("bench_r2j13.json", "bench_radix_r2j13.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "phases", {k: round(v, 3) for k, v in d["phases_ms"].items()},
          "e2e", round(d["e2e"]["ms_per_step"], 2), "repel", round(d["repel"]["ms_per_iter"], 3),
          {k: round(v["ms_per_step"] if "ms_per_step" in v else v["ms_per_iter"], 3) for k, v in d["extras"].items()}, d["parity_check"])
PY
