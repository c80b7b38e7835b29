#!/bin/bash
# quick iteration run: op-level tests, stem / pointwise timings, then the forward parity tests and a short bench
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q --timeout 60 --timeout-method thread 2>&1 | tail -5 | tee gpurun_out/it_ops.log
for args in "--kind stem2 --cin 3 --cout 16 --k 3 --stride 2 --hw 640 --tc 1" \
            "--kind stem2 --cin 3 --cout 32 --k 3 --stride 2 --hw 640 --tc 1 --batch 32" \
            "--kind conv --cin 96 --cout 85 --hw 80 --act 0 --tc 1" \
            "--kind conv --cin 32 --cout 96 --hw 80 --act 0 --up 1 --tc 1" \
            "--kind conv --cin 16 --cout 48 --k 3 --stride 2 --hw 160 --act 1 --tc 2" \
            "--kind conv --cin 48 --cout 96 --hw 40 --act 1 --tc 1" \
            "--kind dwpw --cin 96 --cout 96 --hw 80 --k2 3 --act 1 --tc 1"; do
  timeout 120 python scripts/bench_op.py $args --iters 50 2>&1 | tail -1 | cut -c1-220 | tee -a gpurun_out/it_opbench.log
done
timeout 400 python -m pytest tests/test_gpu_forward.py tests/test_gpu_fullsize.py tests/test_gpu_api.py -x -q --timeout 90 --timeout-method thread 2>&1 | tail -5 | tee gpurun_out/it_fwd.log
timeout 240 python bench.py --no-extra --no-cpu-baseline > gpurun_out/it_bench.json 2> gpurun_out/it_bench.err; tail -c 1500 gpurun_out/it_bench.json
