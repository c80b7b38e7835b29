#!/bin/bash
# stem iteration: stem tests, timings, one ncu capture
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ops.py -x -q -k "stem" --timeout 60 --timeout-method thread 2>&1 | tail -3 | tee gpurun_out/it_ops.log
for args in "--kind stem2 --cin 3 --cout 16 --k 3 --stride 2 --hw 640 --tc 1" \
            "--kind stem2 --cin 3 --cout 32 --k 3 --stride 2 --hw 640 --tc 1 --batch 32"; do
  timeout 120 python scripts/bench_op.py $args --iters 50 2>&1 | tail -1 | cut -c1-220 | tee -a gpurun_out/it_opbench.log
done
timeout 300 bash scripts/gpu_ncu_stem.sh ${1:-stem2_c}
