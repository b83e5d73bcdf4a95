#!/bin/bash
# usage: gpu_prof3.sh <tag> <kernel-regex> <bench args...>: one ncu --set full capture of the kernel (no tests)
TAG=$1; KRE=$2; shift 2
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 4 -c 1 \
   -f -o gpurun_out/prof_$TAG python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" \
   > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"; tail -n 3 gpurun_out/ncu_full_$TAG.log
