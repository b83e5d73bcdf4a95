#!/bin/bash
# One gpurun call: the driver's round-end sequence (GPU parity tests, smoke, both bench arms).  Logs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -q -m gpu -x --tb=short > gpurun_out/test_gpu.log 2>&1
echo "pytest -m gpu exit $?" | tee gpurun_out/summary.txt
tail -n 15 gpurun_out/test_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt
tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_ref.json
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
