#!/bin/bash
set -u
mkdir -p gpurun_out
for b in 16 32 48 64 96 128; do
  timeout 200 python bench.py --no-extra --no-cpu-baseline --no-e2e --batch $b --steps 50 > gpurun_out/sweep_$b.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('gpurun_out/sweep_$b.json').read().strip().splitlines()[-1])
print("batch", $b, "img/s", round(d['value']), "ms", round(d['ms_per_step'],4), "ms/img", round(d['ms_per_step']/$b*1000,2), "us")
PY
done
