#!/bin/bash
# Round 2, GPU job 29 (1 GPU): after restoring radius.cu — shard-only radius on config #4 as every rank of 4, radius parity tests.
out=gpurun_out; mkdir -p $out; tag=r2j29
timeout 300 python scripts/debug_radius_shard.py 10000000 4 > $out/dbg_default_$tag.log 2>&1; echo "default rc=$?"; tail -8 $out/dbg_default_$tag.log
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -x -k "radius or config4 or csr" > $out/pytest_$tag.log 2>&1; echo "pytest_rc=$?" ); tail -3 $out/pytest_$tag.log
