#!/bin/bash
# refresh the launch list and the head-output capture after the last epilogue change
set -u
mkdir -p gpurun_out
R=r02
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-graph > gpurun_out/bench_under_ncu.log 2>&1
KREGEX=tc_conv timeout 200 bash scripts/gpu_ncu_one.sh ${R}_ncu_pw_head_out --kind conv --cin 96 --cout 85 --hw 80 --act 0 --tc 1 > /dev/null 2>&1
python scripts/ncu_top.py gpurun_out/${R}_ncu_pw_head_out.ncu-rep 12 > gpurun_out/${R}_ncu_pw_head_out_summary.txt 2>&1; head -10 gpurun_out/${R}_ncu_pw_head_out_summary.txt | cut -c1-140
