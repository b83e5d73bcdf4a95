#!/bin/bash
# ncu evidence for the bench command: launch list (device time per launch) + one full capture of the
# dominant kernel.  Usage: gpu_profile.sh <tag> <kernel-regex> [bench args...]
TAG=$1; KRE=$2; shift 2
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" \
   > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "ncu launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 2 \
   -f -o gpurun_out/prof_$TAG python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" \
   > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out/ | tail -n 20
