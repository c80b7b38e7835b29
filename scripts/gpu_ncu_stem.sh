#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-stem2_b}
KREGEX=stem2 bash scripts/gpu_ncu_one.sh r02_ncu_$N --kind stem2 --cin 3 --cout 16 --k 3 --stride 2 --hw 640 --tc 1 > /dev/null 2>&1
python scripts/ncu_top.py gpurun_out/r02_ncu_$N.ncu-rep 14 > gpurun_out/r02_ncu_${N}_summary.txt 2>&1; head -30 gpurun_out/r02_ncu_${N}_summary.txt | cut -c1-160
