#!/bin/bash
# final visit: full GPU test suite + smoke + full bench, then the sanitizers on the kernel classes touched last
set -u
bash scripts/gpu_tests.sh
for tool in memcheck racecheck initcheck synccheck; do
  for spec in "head --kind conv --cin 96 --cout 85 --hw 20 --act 0 --tc 2" "head_partial --kind conv --cin 96 --cout 85 --hw 10 --act 0 --tc 2" "stem2 --kind stem2 --cin 3 --cout 16 --k 3 --stride 2 --hw 128 --tc 1"; do
    set -- $spec; name=$1; shift
    timeout 300 compute-sanitizer --tool $tool --print-limit 3 python scripts/bench_op.py "$@" --iters 1 --batch 2 > gpurun_out/r02_san_${tool}_$name.log 2>&1
    echo "[$tool $name] $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02_san_${tool}_$name.log | tail -1)"
  done
done
