#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gmm_int.py -q -m gpu -x --tb=short > gpurun_out/test_int.log 2>&1; echo "int tests exit $?"; tail -n 12 gpurun_out/test_int.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload gmm-int > gpurun_out/bench_gmm-int.json 2> gpurun_out/bench_gmm-int.err; echo "bench exit $?"
cut -c1-330 gpurun_out/bench_gmm-int.json; tail -n 3 gpurun_out/bench_gmm-int.err
