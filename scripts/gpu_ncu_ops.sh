#!/bin/bash
set -u
mkdir -p gpurun_out
prof() { name=$1; shift; timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv -s 3 -c 1 -o gpurun_out/$name -f python scripts/bench_op.py "$@" --iters 3 > gpurun_out/$name.log 2>&1; tail -n 1 gpurun_out/$name.log; }
prof prof_pw96 --kind conv --cin 96 --cout 96 --hw 80 --tc 1
prof prof_up --kind conv --cin 32 --cout 96 --hw 80 --tc 1 --up 1
prof prof_n85 --kind conv --cin 96 --cout 85 --hw 80 --tc 1 --act 0
