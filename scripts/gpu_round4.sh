#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gmm_tensor.py -q -m gpu --tb=short > gpurun_out/test_tensor.log 2>&1; echo "tensor tests exit $?"; tail -n 30 gpurun_out/test_tensor.log
timeout 600 python bench.py --steps 20 --warmup 5 --workload gmm-tensor > gpurun_out/bench_gmm-tensor.json 2> gpurun_out/bench_gmm-tensor.err; echo "bench exit $?"
cat gpurun_out/bench_gmm-tensor.json; tail -n 3 gpurun_out/bench_gmm-tensor.err
bash scripts/gpu_profile.sh gmmtensor gmm_tensor_kernel --workload gmm-tensor
