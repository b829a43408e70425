#!/bin/bash
# Round 2, GPU job 2b (2 GPUs): the multi-GPU tests (one process per GPU: torchrun; one process over both: wtp_create_multi),
# the wall tests whose margins changed, and the bench line at N = 2 with its parity_check.
out=gpurun_out; mkdir -p $out; tag=r2j2b
nvidia-smi -L > $out/host_$tag.txt; nproc >> $out/host_$tag.txt
( timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -k "multi or sharded or mesh_wall or deposit" > $out/pytest_$tag.log 2>&1; echo "pytest_rc=$?" )
tail -15 $out/pytest_$tag.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err
echo "bench_rc=$?"; cut -c1-300 $out/bench_$tag.json; tail -5 $out/bench_$tag.err
