#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_pipeline.py -q -m gpu --tb=short > gpurun_out/test_fe.log 2>&1; echo "frontend tests exit $?"; tail -n 25 gpurun_out/test_fe.log
for wl in "$@"; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl exit $?"
cut -c1-330 gpurun_out/bench_$wl.json; tail -n 3 gpurun_out/bench_$wl.err
done
