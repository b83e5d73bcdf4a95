#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_nn.py -q -m gpu --tb=short -x > gpurun_out/test_nn.log 2>&1; echo "nn tests exit $?"; tail -n 8 gpurun_out/test_nn.log
for v in 0 1; do
if [ $v = 1 ]; then export RB_NN_NO_TMA_STORE=1; fi
timeout 300 python bench.py --steps 10 --warmup 3 --workload nn --frames 75776 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('no_tma_store $v', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'])"
done
