#!/bin/bash
# Round 2, GPU job 2a (1 GPU): the whole -m gpu suite on the current build, then the default bench line.
out=gpurun_out; mkdir -p $out; tag=r2j2a
( timeout 2400 python -m pytest tests -m gpu -q > $out/pytest_$tag.log 2>&1; echo "pytest_rc=$?" )
tail -6 $out/pytest_$tag.log
timeout 900 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench_rc=$?"; cut -c1-300 $out/bench_$tag.json; tail -3 $out/bench_$tag.err
