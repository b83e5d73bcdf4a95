#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_search.py -q -m gpu --tb=short > gpurun_out/test_search.log 2>&1; echo "search tests exit $?"; tail -n 25 gpurun_out/test_search.log
grep search_c5 gpurun_out/diag.jsonl | tail -1
