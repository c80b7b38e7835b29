#!/bin/bash
# quick regression visit: op-level + forward + API parity tests, a few op timings, the headline bench (no extras)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q --timeout 60 --timeout-method thread 2>&1 | tail -3 | tee gpurun_out/q_ops.log
rm -f gpurun_out/q_opbench.log
for args in "--kind stem2 --cin 3 --cout 16 --k 3 --stride 2 --hw 640 --tc 1" \
            "--kind conv --cin 96 --cout 85 --hw 80 --act 0 --tc 1" \
            "--kind conv --cin 96 --cout 85 --hw 40 --act 0 --tc 1" \
            "--kind conv --cin 32 --cout 96 --hw 80 --act 0 --up 1 --tc 1" \
            "--kind dwpw --cin 96 --cout 96 --hw 80 --k2 3 --act 1 --tc 1" \
            "--kind dwpw --cin 256 --cout 64 --hw 20 --k2 5 --act 0 --act2 1 --res 1 --tc 1" \
            "--kind dwpw --cin 96 --cout 48 --hw 40 --k2 3 --act 0 --act2 1 --res 1 --tc 1"; do
  timeout 120 python scripts/bench_op.py $args --iters 50 2>&1 | tail -1 | cut -c1-30,100-230 | tee -a gpurun_out/q_opbench.log
done
timeout 400 python -m pytest tests/test_gpu_forward.py tests/test_gpu_fullsize.py tests/test_gpu_api.py -x -q --timeout 90 --timeout-method thread 2>&1 | tail -3 | tee gpurun_out/q_fwd.log
timeout 240 python bench.py --no-extra --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/q_bench.json').read().strip().splitlines()[-1])
print("bench value", round(d['value']), "ms", round(d['ms_per_step'],4), "e2e", round(d['e2e']['value']), "stem frac", round(d['roofline']['frac'],4), "stem ms", round(d['roofline']['kernel_ms'],4))
PY
