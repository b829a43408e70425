#!/bin/bash
# Round 2, GPU job 26 (1 GPU): radius fill with the software-pipelined long-row merge — radius parity tests + timing.
out=gpurun_out; mkdir -p $out; tag=${1:-r2j26}
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py tests/test_gpu_index_build.py -m gpu -q -x -k "radius or config4 or csr" > $out/pytest_$tag.log 2>&1; echo "pytest_rc=$?" ); tail -4 $out/pytest_$tag.log
timeout 600 python scripts/radius_time.py > $out/radius_time_$tag.log 2>&1; echo "time_rc=$?"; cat $out/radius_time_$tag.log | tail -4
