#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 100 --warmup 10 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "rc=$?"; head -c 200 gpurun_out/bench_n4.json; echo; tail -2 gpurun_out/bench_n4.err
