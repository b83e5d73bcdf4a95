#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nn.py -q -m gpu --tb=short > gpurun_out/test_nn.log 2>&1; echo "nn tests exit $?"; tail -n 15 gpurun_out/test_nn.log
timeout 300 python scripts/gemm_bench.py > gpurun_out/gemm_bench.jsonl 2>&1; cat gpurun_out/gemm_bench.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --workload nn --frames 75776 > gpurun_out/bench_nn.json 2> gpurun_out/bench_nn.err; echo "bench exit $?"
cat gpurun_out/bench_nn.json; tail -n 3 gpurun_out/bench_nn.err
