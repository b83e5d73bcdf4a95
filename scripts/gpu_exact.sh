#!/bin/bash
# the exact two-pass batch-float scorer: parity tests, then the C2 bench line and a launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gmm_exact.py tests/test_gpu_gmm.py tests/test_gpu_gmm_tensor.py -q -x --tb=short > gpurun_out/test_exact.log 2>&1
echo "exact tests exit $?" | tee gpurun_out/summary_exact.txt
tail -n 40 gpurun_out/test_exact.log
timeout 300 python bench.py --workload gmm --no-cpu-baseline > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err
echo "bench exit $?" | tee -a gpurun_out/summary_exact.txt
cat gpurun_out/bench_exact.json | cut -c1-1500; tail -n 5 gpurun_out/bench_exact.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_exact.csv \
    python bench.py --workload gmm --no-cpu-baseline --steps 3 --warmup 3 > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_exact.csv | awk -F'","' '{print $5, $(NF)}' | tail -n 12
