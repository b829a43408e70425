#!/bin/bash
# Round 2, GPU job 36 (1 GPU): build choice by size — 100 M points (radix build above 64 MB of order array), 16 M points either way.
out=gpurun_out; mkdir -p $out
timeout 200 python scripts/check_100M.py 100000000 > $out/check_100M_r2j36.log 2>&1; echo "rc=$?"; tail -2 $out/check_100M_r2j36.log
timeout 100 python scripts/check_100M.py 16000000 > $out/check_16M_r2j36.log 2>&1; echo "rc=$?"; tail -2 $out/check_16M_r2j36.log
WTP_RADIX_BUILD=1 timeout 100 python scripts/check_100M.py 16000000 > $out/check_16M_radix_r2j36.log 2>&1; echo "rc=$?"; tail -2 $out/check_16M_radix_r2j36.log
