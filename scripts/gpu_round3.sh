#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gmm_tensor.py tests/test_gpu_nn.py -q -m gpu --tb=short > gpurun_out/test_tensor.log 2>&1; echo "tensor tests exit $?"; tail -n 30 gpurun_out/test_tensor.log
for wl in gmm-tensor; do
  timeout 600 python bench.py --steps 20 --warmup 5 --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl exit $?"
  cat gpurun_out/bench_$wl.json; tail -n 3 gpurun_out/bench_$wl.err
done
bash scripts/gpu_profile.sh gmmtensor gemm16_kernel --workload gmm-tensor
bash scripts/gpu_profile.sh nn gemm16_kernel --workload nn --frames 16384
bash scripts/gpu_profile.sh frontend mfcc_static --workload frontend
