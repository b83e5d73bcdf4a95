#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gmm.py tests/test_gpu_pipeline.py -q -m gpu -x --tb=short > gpurun_out/test_gmm.log 2>&1; echo "gmm tests exit $?"; tail -n 3 gpurun_out/test_gmm.log
for fpt in 1 2 4; do
RB_GMM_FPT=$fpt timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload gmm 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('fpt $fpt', d['value'], d['ms_per_step'], d['roofline']['fp32_alu']['frac'])"
done
