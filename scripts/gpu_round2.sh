#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_frontend.py -q -m gpu --tb=short > gpurun_out/test_frontend.log 2>&1; echo "frontend exit $?"; tail -n 5 gpurun_out/test_frontend.log
for wl in gmm gmm-diag frontend pipeline nn; do
  timeout 600 python bench.py --steps 20 --warmup 5 --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl exit $?"
  cat gpurun_out/bench_$wl.json; tail -n 3 gpurun_out/bench_$wl.err
done
bash scripts/gpu_profile.sh gmm gmm_batch_kernel
