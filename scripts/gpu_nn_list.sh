#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_nn.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload nn --frames 37888 > gpurun_out/ncu_nn.log 2>&1
echo "exit $?"
