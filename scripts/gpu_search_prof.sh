#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:linear_search --csv --log-file gpurun_out/launches_search.csv python -m pytest tests/test_gpu_search.py -q -m gpu -k "scores_from_the_gmm" > gpurun_out/ncu_search.log 2>&1
echo "exit $?"
