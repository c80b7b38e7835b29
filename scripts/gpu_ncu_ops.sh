#!/bin/bash
# ncu --set full of one representative launch per kernel class (single-op micro-benchmarks, batch 64); summaries -> gpurun_out/
set -u
mkdir -p gpurun_out
cap() { name=$1; shift; timeout 200 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-tc_conv}" -s 3 -c 1 -o gpurun_out/$name -f python scripts/bench_op.py "$@" --iters 3 > gpurun_out/$name.log 2>&1; python scripts/ncu_top.py gpurun_out/$name.ncu-rep 6 > gpurun_out/${name}_summary.txt 2>&1; head -9 gpurun_out/${name}_summary.txt | sed "s/^/$name: /"; }
cap ncu_dwpw_p3 --kind dwpw --cin 96 --cout 96 --hw 80 --tc 1
cap ncu_dwpw_uir_res --kind dwpw --cin 96 --cout 48 --hw 40 --tc 1 --k2 3 --act2 1 --act 0 --res 1
cap ncu_pw_head_out --kind conv --cin 96 --cout 85 --hw 80 --tc 1 --act 0
cap ncu_pw_lateral_up --kind conv --cin 32 --cout 96 --hw 80 --tc 1 --up 1 --act 0
cap ncu_conv3x3s2 --kind conv --cin 16 --cout 48 --k 3 --stride 2 --hw 160 --tc 1
rm -f gpurun_out/ncu_dwpw_p3.ncu-rep gpurun_out/ncu_dwpw_uir_res.ncu-rep gpurun_out/ncu_pw_head_out.ncu-rep gpurun_out/ncu_pw_lateral_up.ncu-rep gpurun_out/ncu_conv3x3s2.ncu-rep
