#!/bin/bash
# ncu --set full of every kernel class (one launch each): gpurun_out/${R}_ncu_<class>.ncu-rep + _summary.txt
set -u
mkdir -p gpurun_out
R=${1:-r02b}
cap() { # name kregex bench_op-args...
  local name=$1 kre=$2; shift 2
  KREGEX=$kre timeout 200 bash scripts/gpu_ncu_one.sh ${R}_ncu_$name "$@" > /dev/null 2>&1
  python scripts/ncu_top.py gpurun_out/${R}_ncu_$name.ncu-rep 12 > gpurun_out/${R}_ncu_${name}_summary.txt 2>&1
  head -10 gpurun_out/${R}_ncu_${name}_summary.txt | sed "s/^/[$name] /" | cut -c1-140
}
cap stem2 stem2 --kind stem2 --cin 3 --cout 16 --k 3 --stride 2 --hw 640 --tc 1
cap pw_lateral_up tc_conv --kind conv --cin 32 --cout 96 --hw 80 --act 0 --up 1 --tc 1
cap pw_head_out tc_conv --kind conv --cin 96 --cout 85 --hw 80 --act 0 --tc 1
cap dwpw_k3_p3 tc_conv --kind dwpw --cin 96 --cout 96 --hw 80 --k2 3 --act 1 --tc 1
cap dwpw_k5_uir tc_conv --kind dwpw --cin 256 --cout 64 --hw 20 --k2 5 --act 0 --act2 1 --res 1 --tc 1
cap dwpw_k3_s2 tc_conv --kind dwpw --cin 288 --cout 64 --hw 40 --k2 3 --stride2 2 --act 0 --act2 1 --tc 1
cap conv3x3_s2 tc_conv --kind conv --cin 16 --cout 48 --k 3 --stride 2 --hw 160 --act 1 --tc 2
cap dense3x3_tap tc_conv --kind conv --cin 328 --cout 328 --k 3 --hw 160 --batch 4 --act 2 --tc 2 --tap 1
cap pw_small tc_conv --kind conv --cin 64 --cout 256 --hw 20 --act 1 --tc 1
