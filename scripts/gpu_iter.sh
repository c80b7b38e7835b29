#!/bin/bash
# quick iteration: op numerics + op micro-benchmarks (+ optional env for A/B)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q 2>&1 | tail -5
echo "== RAWHI=0 numerics"; YL_TC_RAWHI=0 timeout 600 python -m pytest tests/test_gpu_ops.py -x -q 2>&1 | tail -3 | cut -c1-200
for raw in 1; do
echo "== RAWHI=$raw"
YL_TC_RAWHI=$raw python scripts/bench_op.py --kind conv --cin 96 --cout 96 --hw 80 --tc 1
YL_TC_RAWHI=$raw python scripts/bench_op.py --kind conv --cin 32 --cout 96 --hw 80 --tc 1
YL_TC_RAWHI=$raw python scripts/bench_op.py --kind conv --cin 32 --cout 96 --hw 80 --tc 1 --up 1
YL_TC_RAWHI=$raw python scripts/bench_op.py --kind conv --cin 96 --cout 48 --hw 40 --tc 1 --res 1 --act 0
YL_TC_RAWHI=$raw python scripts/bench_op.py --kind dwpw --cin 96 --cout 96 --hw 80 --tc 1
YL_TC_RAWHI=$raw python scripts/bench_op.py --kind dwpw --cin 96 --cout 96 --hw 40 --tc 1
YL_TC_RAWHI=$raw python scripts/bench_op.py --kind dwpw --cin 96 --cout 96 --hw 20 --tc 1
YL_TC_RAWHI=$raw python scripts/bench_op.py --kind conv --cin 16 --cout 48 --k 3 --stride 2 --hw 160 --tc 1
YL_TC_RAWHI=$raw python scripts/bench_op.py --kind stem2 --cin 3 --cout 16 --k 3 --stride 2 --hw 640 --tc 1
YL_TC_RAWHI=$raw python scripts/bench_op.py --kind conv --cin 96 --cout 85 --hw 80 --tc 1 --act 0
done 2>&1 | tee gpurun_out/bench_ops.log
