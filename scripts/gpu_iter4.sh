#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/it_opbench.log
for cfg in 0 1; do
  echo "== YL_STEM_CFG=$cfg" | tee -a gpurun_out/it_opbench.log
  YL_STEM_CFG=$cfg timeout 200 python -m pytest tests/test_gpu_ops.py -x -q -k "stem" --timeout 60 --timeout-method thread 2>&1 | tail -1 | tee -a gpurun_out/it_opbench.log
  for args in "--kind stem2 --cin 3 --cout 16 --k 3 --stride 2 --hw 640 --tc 1" \
              "--kind stem2 --cin 3 --cout 32 --k 3 --stride 2 --hw 640 --tc 1 --batch 32"; do
    YL_STEM_CFG=$cfg timeout 120 python scripts/bench_op.py $args --iters 50 2>&1 | tail -1 | cut -c100-220 | tee -a gpurun_out/it_opbench.log
  done
done
timeout 300 python -m pytest tests/test_gpu_forward.py tests/test_gpu_api.py -x -q --timeout 90 --timeout-method thread 2>&1 | tail -3 | tee gpurun_out/it_fwd.log
timeout 240 python bench.py --no-extra --no-cpu-baseline > gpurun_out/it_bench.json 2> gpurun_out/it_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/it_bench.json').read().strip().splitlines()[-1])
print("bench", d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('kernel_ms'))
PY
YL_STEM_CFG=1 timeout 240 python bench.py --no-extra --no-cpu-baseline --no-e2e > gpurun_out/it_bench1.json 2> gpurun_out/it_bench1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/it_bench1.json').read().strip().splitlines()[-1])
print("bench cfg1", d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('kernel_ms'))
PY
