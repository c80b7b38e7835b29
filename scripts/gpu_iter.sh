#!/bin/bash
# quick iteration on a B200 box: op numerics + forward parity + a handful of single-op micro-benchmarks + the bench line
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py -q 2>&1 | grep -E "FAILED|passed|failed" | head
{
python scripts/bench_op.py --kind stem2 --cin 3 --cout 16 --k 3 --stride 2 --hw 640 --tc 1
python scripts/bench_op.py --kind conv --cin 16 --cout 48 --k 3 --stride 2 --hw 160 --tc 1
python scripts/bench_op.py --kind dwpw --cin 32 --cout 96 --hw 80 --tc 1 --k2 5
python scripts/bench_op.py --kind dwpw --cin 96 --cout 48 --hw 80 --tc 1 --k2 5 --act2 1 --act 0 --stride2 2
python scripts/bench_op.py --kind dwpw --cin 96 --cout 48 --hw 40 --tc 1 --k2 3 --act2 1 --act 0 --res 1
python scripts/bench_op.py --kind dwpw --cin 256 --cout 64 --hw 20 --tc 1 --k2 5 --act2 1 --act 0 --res 1
python scripts/bench_op.py --kind conv --cin 48 --cout 96 --hw 40 --tc 1
python scripts/bench_op.py --kind conv --cin 480 --cout 96 --hw 20 --tc 1 --act 0
python scripts/bench_op.py --kind conv --cin 32 --cout 96 --hw 80 --tc 1 --up 1 --act 0
python scripts/bench_op.py --kind dwpw --cin 96 --cout 96 --hw 80 --tc 1
python scripts/bench_op.py --kind conv --cin 96 --cout 85 --hw 80 --tc 1 --act 0
} 2>&1 | grep ms | tee gpurun_out/bench_ops.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dump-ops gpurun_out/op_times.json 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-200
tail -3 gpurun_out/bench.err
