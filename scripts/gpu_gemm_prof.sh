#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm16_mt2 -s 2 -c 1 -f -o gpurun_out/prof_gemm_mt2 python scripts/gemm_bench.py > gpurun_out/ncu_gemm.log 2>&1
echo "exit $?"; tail -n 3 gpurun_out/ncu_gemm.log
