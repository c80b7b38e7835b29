#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_post.py tests/test_gpu_api.py -x -q --timeout 90 --timeout-method thread 2>&1 | tail -3 | tee gpurun_out/q_post.log
timeout 120 python scripts/post_bench.py 2>&1 | tail -2
timeout 240 python bench.py --no-extra --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/q_bench.json').read().strip().splitlines()[-1])
print("bench value", round(d['value']), "ms", round(d['ms_per_step'],4), "e2e", round(d['e2e']['value']), "post ms", d['roofline_step'].get('post_ms'))
PY
