#!/bin/bash
# GMM parity tests + C2 bench (no CPU baseline)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gmm.py tests/test_gpu_pipeline.py -q -m gpu -x --tb=short > gpurun_out/test_gmm.log 2>&1; echo "gmm tests exit $?"; tail -n 5 gpurun_out/test_gmm.log
for wl in "$@"; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl exit $?"
cat gpurun_out/bench_$wl.json; tail -n 3 gpurun_out/bench_$wl.err
done
