#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "rc=$?"; head -c 300 gpurun_out/bench_n2.json; echo; tail -3 gpurun_out/bench_n2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-600
