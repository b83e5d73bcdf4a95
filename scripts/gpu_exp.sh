#!/bin/bash
for lib in "$@"; do
RASR_B200_LIB=$PWD/rasr_b200/lib/exp/$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload gmm 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$lib', d['value'], d['ms_per_step'])"
done
