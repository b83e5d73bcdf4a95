#!/bin/bash
# One gpurun call: GPU parity tests (per file, each under its own timeout), smoke, bench.  Logs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
for f in gmm frontend pipeline nn; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu -x --tb=short > gpurun_out/test_$f.log 2>&1
  echo "test_gpu_$f exit $?" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/test_$f.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
