#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_search.py -q -m gpu --tb=short > gpurun_out/test_search.log 2>&1; echo "search tests exit $?"; tail -n 25 gpurun_out/test_search.log
grep search_c5 gpurun_out/diag.jsonl | tail -1
timeout 300 python bench.py --workload pipeline-search --steps 5 --warmup 3 | tee gpurun_out/bench_pipeline_search.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:linear_search --csv --log-file gpurun_out/launches_search.csv python -m pytest tests/test_gpu_search.py -q -m gpu -k "scores_from_the_gmm" > gpurun_out/ncu_search.log 2>&1
grep linear_search gpurun_out/launches_search.csv | tail -3
