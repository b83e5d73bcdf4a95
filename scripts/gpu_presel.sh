#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gmm_presel.py tests/test_gpu_gmm.py -q -m gpu --tb=short -x > gpurun_out/test_presel.log 2>&1; echo "presel tests exit $?"; tail -n 30 gpurun_out/test_presel.log
grep presel gpurun_out/diag.jsonl | tail -3
