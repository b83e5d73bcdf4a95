#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_postproc.py -q -m gpu --tb=short > gpurun_out/test_pp.log 2>&1; echo "postproc tests exit $?"; tail -n 25 gpurun_out/test_pp.log
grep postproc gpurun_out/diag.jsonl | tail -2
