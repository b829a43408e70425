#!/bin/bash
# Round 2, GPU job 33 (1 GPU): the whole -m gpu suite, smoke() and the default bench line on the final tree.
out=gpurun_out; mkdir -p $out; tag=r2j33
( timeout 2400 python -m pytest tests -m gpu -q > $out/pytest_$tag.log 2>&1; echo "pytest_rc=$?" ); tail -4 $out/pytest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench_rc=$?"; cut -c1-200 $out/bench_$tag.json
