#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nn.py -q -m gpu --tb=short > gpurun_out/test_nn.log 2>&1; echo "nn tests exit $?"; tail -n 5 gpurun_out/test_nn.log
timeout 300 python scripts/gemm_bench.py 2>&1 
timeout 600 python bench.py --steps 10 --warmup 3 --workload nn --frames 75776 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
