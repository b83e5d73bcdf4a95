#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --tb=short > gpurun_out/test_pipe.log 2>&1; echo "pipeline tests exit $?"; tail -n 12 gpurun_out/test_pipe.log
timeout 600 python bench.py --steps 10 --warmup 3 --workload pipeline-nn --no-cpu-baseline > gpurun_out/bench_pipeline-nn.json 2> gpurun_out/bench_pipeline-nn.err; echo "bench exit $?"; tail -n1 gpurun_out/bench_pipeline-nn.json | cut -c1-900; tail -n 3 gpurun_out/bench_pipeline-nn.err
