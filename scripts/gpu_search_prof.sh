#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:linear_search -c 1 -o gpurun_out/search_full -f python -m pytest tests/test_gpu_search.py -q -m gpu -k "scores_from_the_gmm" > gpurun_out/ncu_search_full.log 2>&1
echo "exit $?"; tail -3 gpurun_out/ncu_search_full.log
