#!/bin/bash
# usage: gpu_ncu_one.sh <name> <bench_op args...>
set -u
mkdir -p gpurun_out
name=$1; shift
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-tc_conv}" -s 3 -c 1 -o gpurun_out/$name -f python scripts/bench_op.py "$@" --iters 3 > gpurun_out/$name.log 2>&1
tail -n 1 gpurun_out/$name.log
