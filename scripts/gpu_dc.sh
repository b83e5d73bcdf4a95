#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_frontend_dc.py tests/test_gpu_frontend.py tests/test_gpu_pipeline.py -q -m gpu --tb=short -x > gpurun_out/test_dc.log 2>&1; echo "dc tests exit $?"; tail -n 30 gpurun_out/test_dc.log
