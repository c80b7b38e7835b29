#!/bin/bash
# Round check on a B200 box: parity tests, smoke, bench (both arms), ncu launch list, ncu --set full of the top kernel.
set -u
mkdir -p gpurun_out
bash scripts/gpu_check.sh
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | tee gpurun_out/bench_ref.json
echo "== ncu full: fused stem kernel"
KREGEX=stem2 bash scripts/gpu_ncu_one.sh ncu_stem2 --kind stem2 --cin 3 --cout 16 --k 3 --stride 2 --hw 640 --tc 1
python scripts/ncu_top.py gpurun_out/ncu_stem2.ncu-rep 30 > gpurun_out/ncu_stem2_summary.txt 2>&1
tail -45 gpurun_out/ncu_stem2_summary.txt
