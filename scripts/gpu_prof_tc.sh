#!/bin/bash
set -u
mkdir -p gpurun_out
for tc in 0 1; do
python scripts/bench_op.py --kind conv --cin 96 --cout 96 --hw 80 --tc $tc
python scripts/bench_op.py --kind conv --cin 32 --cout 96 --hw 80 --tc $tc
python scripts/bench_op.py --kind conv --cin 32 --cout 96 --hw 80 --tc $tc --up 1
python scripts/bench_op.py --kind dwpw --cin 96 --cout 96 --hw 80 --tc $tc
python scripts/bench_op.py --kind conv --cin 32 --cout 16 --k 3 --stride 2 --hw 320 --tc $tc
python scripts/bench_op.py --kind conv --cin 16 --cout 16 --hw 160 --tc $tc
done 2>&1 | tee gpurun_out/bench_ops.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv -s 3 -c 1 -o gpurun_out/prof_tc_pw96 -f \
   python scripts/bench_op.py --kind conv --cin 96 --cout 96 --hw 80 --tc 1 --iters 3 > gpurun_out/ncu_pw.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv -s 3 -c 1 -o gpurun_out/prof_tc_dwpw96 -f \
   python scripts/bench_op.py --kind dwpw --cin 96 --cout 96 --hw 80 --tc 1 --iters 3 > gpurun_out/ncu_dwpw.log 2>&1
tail -2 gpurun_out/ncu_pw.log gpurun_out/ncu_dwpw.log
