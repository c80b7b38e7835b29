#!/bin/bash
# Round-2 evidence run on a B200 box: ncu launch list of the bench step, ncu --set full captures of every kernel class, and
# compute-sanitizer (memcheck / racecheck / initcheck / synccheck) on one launch of every kernel class.  Output: gpurun_out/.
set -u
mkdir -p gpurun_out
R=r02
echo "== ncu launch list (2 steps of the bench workload, eager launches so that every kernel is listed)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-graph > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-200
cap() { # name kregex bench_op-args...
  local name=$1 kre=$2; shift 2
  KREGEX=$kre bash scripts/gpu_ncu_one.sh ${R}_ncu_$name "$@" > /dev/null 2>&1
  python scripts/ncu_top.py gpurun_out/${R}_ncu_$name.ncu-rep 12 > gpurun_out/${R}_ncu_${name}_summary.txt 2>&1
  head -12 gpurun_out/${R}_ncu_${name}_summary.txt | sed "s/^/[$name] /" | cut -c1-140
}
echo "== ncu --set full per kernel class"
cap stem2 stem2 --kind stem2 --cin 3 --cout 16 --k 3 --stride 2 --hw 640 --tc 1
cap pw_lateral_up tc_conv --kind conv --cin 32 --cout 96 --hw 80 --act 0 --up 1 --tc 1
cap pw_head_out tc_conv --kind conv --cin 96 --cout 85 --hw 80 --act 0 --tc 1
cap dwpw_k3_p3 tc_conv --kind dwpw --cin 96 --cout 96 --hw 80 --k2 3 --act 1 --tc 1
cap dwpw_k5_uir tc_conv --kind dwpw --cin 256 --cout 64 --hw 20 --k2 5 --act 0 --act2 1 --res 1 --tc 1
cap dwpw_k3_s2 tc_conv --kind dwpw --cin 288 --cout 64 --hw 40 --k2 3 --stride2 2 --act 0 --act2 1 --tc 1
cap conv3x3_s2 tc_conv --kind conv --cin 16 --cout 48 --k 3 --stride 2 --hw 160 --act 1 --tc 2
cap dense3x3_tap tc_conv --kind conv --cin 328 --cout 328 --k 3 --hw 160 --batch 4 --act 2 --tc 2 --tap 1
cap simt_pw conv_gemm --kind conv --cin 16 --cout 16 --hw 160 --act 1 --tc 0
cap simt_dw dw_kernel --kind dw --cin 96 --cout 96 --k 3 --hw 80 --act 0 --tc 0
KREGEX=post_kernel timeout 300 ncu --set full --clock-control none --import-source on -k regex:post_kernel -s 2 -c 1 -o gpurun_out/${R}_ncu_post -f python scripts/post_bench.py > gpurun_out/${R}_ncu_post.log 2>&1
python scripts/ncu_top.py gpurun_out/${R}_ncu_post.ncu-rep 12 > gpurun_out/${R}_ncu_post_summary.txt 2>&1; head -12 gpurun_out/${R}_ncu_post_summary.txt | sed "s/^/[post] /" | cut -c1-140
KREGEX=pre_kernel timeout 300 ncu --set full --clock-control none --import-source on -k regex:pre_kernel -s 1 -c 1 -o gpurun_out/${R}_ncu_pre -f python scripts/pre_bench.py > gpurun_out/${R}_ncu_pre.log 2>&1
python scripts/ncu_top.py gpurun_out/${R}_ncu_pre.ncu-rep 12 > gpurun_out/${R}_ncu_pre_summary.txt 2>&1; head -12 gpurun_out/${R}_ncu_pre_summary.txt | sed "s/^/[pre] /" | cut -c1-140
echo "== compute-sanitizer"
san() { # tool name args...
  local tool=$1 name=$2; shift 2
  timeout 600 compute-sanitizer --tool $tool --print-limit 3 python scripts/bench_op.py "$@" --iters 1 --batch 2 > gpurun_out/${R}_san_${tool}_$name.log 2>&1
  echo "[$tool $name] $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${R}_san_${tool}_$name.log | tail -1)"
}
for tool in memcheck racecheck initcheck synccheck; do
  san $tool stem2 --kind stem2 --cin 3 --cout 16 --k 3 --stride 2 --hw 128 --tc 1
  san $tool pw --kind conv --cin 48 --cout 96 --hw 20 --act 1 --tc 2
  san $tool pw_up --kind conv --cin 32 --cout 96 --hw 20 --act 0 --up 1 --tc 2
  san $tool head --kind conv --cin 96 --cout 85 --hw 20 --act 0 --tc 2
  san $tool dwpw3 --kind dwpw --cin 96 --cout 96 --hw 20 --k2 3 --act 1 --tc 2
  san $tool dwpw5res --kind dwpw --cin 256 --cout 64 --hw 20 --k2 5 --act 0 --act2 1 --res 1 --tc 2
  san $tool dwpw3s2 --kind dwpw --cin 288 --cout 64 --hw 40 --k2 3 --stride2 2 --act 0 --act2 1 --tc 2
  san $tool conv3x3s2 --kind conv --cin 16 --cout 48 --k 3 --stride 2 --hw 40 --act 1 --tc 2
  san $tool dense3x3tap --kind conv --cin 196 --cout 196 --k 3 --hw 20 --act 2 --tc 2 --tap 1
done
timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python scripts/post_bench.py > gpurun_out/${R}_san_memcheck_post.log 2>&1; echo "[memcheck post] $(grep 'ERROR SUMMARY' gpurun_out/${R}_san_memcheck_post.log | tail -1)"
timeout 600 compute-sanitizer --tool racecheck --print-limit 3 python scripts/post_bench.py > gpurun_out/${R}_san_racecheck_post.log 2>&1; echo "[racecheck post] $(grep 'RACECHECK SUMMARY' gpurun_out/${R}_san_racecheck_post.log | tail -1)"
