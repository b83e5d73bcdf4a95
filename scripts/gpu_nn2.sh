#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gemm_bench.py 2>&1 | grep 'variant'
for db in 0 1; do export RB_NN_MT2=$db;
timeout 600 python bench.py --steps 10 --warmup 3 --workload nn --frames 75776 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('final_db $db', d['value'], d['ms_per_step'], d['roofline']['achieved'])"
done
RB_NN_MT2=0 timeout 600 python -m pytest tests/test_gpu_nn.py -q -m gpu --tb=short 2>&1 | tail -2
