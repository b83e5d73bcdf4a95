#!/bin/bash
# usage: gpu_prof2.sh <tag> <kernel-regex> <bench args...>: launch list + one ncu --set full capture
TAG=$1; KRE=$2; shift 2
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
   --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" \
   > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "ncu launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 4 -c 1 \
   -f -o gpurun_out/prof_$TAG python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" \
   > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"; tail -n 2 gpurun_out/ncu_full_$TAG.log | cut -c1-200
