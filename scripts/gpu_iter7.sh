#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_forward.py -x -q 2>&1 | tail -4
{
for t in 0 1; do
echo "YL_TC_TMAOUT=$t"
export YL_TC_TMAOUT=$t
python scripts/bench_op.py --kind dwpw --cin 96 --cout 96 --hw 80 --tc 1
python scripts/bench_op.py --kind dwpw --cin 96 --cout 96 --hw 40 --tc 1
python scripts/bench_op.py --kind dwpw --cin 32 --cout 96 --hw 80 --tc 1 --k2 5
python scripts/bench_op.py --kind conv --cin 48 --cout 96 --hw 40 --tc 1
python scripts/bench_op.py --kind conv --cin 64 --cout 256 --hw 20 --tc 1
python scripts/bench_op.py --kind conv --cin 48 --cout 32 --hw 80 --tc 1
python scripts/bench_op.py --kind conv --cin 16 --cout 48 --k 3 --stride 2 --hw 160 --tc 1
done
unset YL_TC_TMAOUT
} 2>&1 | tee gpurun_out/bench_ops.log
echo "== bench" ; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dump-ops gpurun_out/op_times.json 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-200
tail -5 gpurun_out/bench.err
