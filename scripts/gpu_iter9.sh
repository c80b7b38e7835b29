#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|passed|failed" | head
{
python scripts/bench_op.py --kind dwpw --cin 96 --cout 48 --hw 40 --tc 1 --k2 3 --act2 1 --act 0 --res 1
python scripts/bench_op.py --kind dwpw --cin 256 --cout 64 --hw 20 --tc 1 --k2 5 --act2 1 --act 0 --res 1
python scripts/bench_op.py --kind conv --cin 192 --cout 48 --hw 40 --tc 1 --res 1 --act 0
python scripts/bench_op.py --kind conv --cin 32 --cout 96 --hw 80 --tc 1 --up 1 --act 0
python scripts/bench_op.py --kind dwpw --cin 96 --cout 96 --hw 80 --tc 1
} 2>&1 | grep ms | cut -c1-40,70-200 | tee gpurun_out/bench_ops.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dump-ops gpurun_out/op_times.json 2>gpurun_out/bench.err | tee gpurun_out/bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
tail -3 gpurun_out/bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/op_times.json'))
print(' '.join(f"{o['i']}:{o['kind'].split('<')[-1][:5]}{o.get('cin','')}-{o.get('cout','')}:{o['ms']*1000:.0f}" for o in d))
P
