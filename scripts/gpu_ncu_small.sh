#!/bin/bash
# ncu --set full captures of the small-map kernel classes + the fused stem (one launch each), summaries into gpurun_out/
set -u
mkdir -p gpurun_out
KREGEX=tc_conv bash scripts/gpu_ncu_one.sh ncu_dwpw_k5_256_64_s20 --kind dwpw --cin 256 --cout 64 --hw 20 --k2 5 --act 0 --act2 1 --res 1 --tc 1
KREGEX=tc_conv bash scripts/gpu_ncu_one.sh ncu_pw_64_256_s20 --kind conv --cin 64 --cout 256 --hw 20 --act 1 --tc 1
KREGEX=tc_conv bash scripts/gpu_ncu_one.sh ncu_dwpw_k3_96_96_s80 --kind dwpw --cin 96 --cout 96 --hw 80 --k2 3 --act 1 --tc 1
KREGEX=stem2 bash scripts/gpu_ncu_one.sh ncu_stem2 --kind stem2 --cin 3 --cout 16 --k 3 --stride 2 --hw 640 --tc 1
for n in ncu_dwpw_k5_256_64_s20 ncu_pw_64_256_s20 ncu_dwpw_k3_96_96_s80 ncu_stem2; do
  python scripts/ncu_top.py gpurun_out/$n.ncu-rep 25 > gpurun_out/${n}_summary.txt 2>&1
done
