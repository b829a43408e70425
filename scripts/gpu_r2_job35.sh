#!/bin/bash
# Round 2, GPU job 35 (1 GPU): the size of BASELINE config #5 (100 M points) on one GPU with the final build.
out=gpurun_out; mkdir -p $out
timeout 280 python scripts/check_100M.py 100000000 > $out/check_100M_r2j35.log 2>&1; echo "rc=$?"; tail -5 $out/check_100M_r2j35.log
