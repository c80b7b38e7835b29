#!/bin/bash
# refresh the launch list and the postprocess capture after the last change, then the full test + bench visit
set -u
mkdir -p gpurun_out
R=r02
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-graph > gpurun_out/bench_under_ncu.log 2>&1
KREGEX=post_kernel timeout 300 ncu --set full --clock-control none --import-source on -k regex:post_kernel -s 2 -c 1 -o gpurun_out/${R}_ncu_post -f python scripts/post_bench.py > gpurun_out/${R}_ncu_post.log 2>&1
python scripts/ncu_top.py gpurun_out/${R}_ncu_post.ncu-rep 12 > gpurun_out/${R}_ncu_post_summary.txt 2>&1; head -4 gpurun_out/${R}_ncu_post_summary.txt | cut -c1-140
bash scripts/gpu_tests.sh
