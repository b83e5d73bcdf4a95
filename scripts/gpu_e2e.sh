#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --tb=short > gpurun_out/test_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -n 4 gpurun_out/test_gpu.log
for wl in frontend pipeline pipeline-search gmm; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl exit $?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$wl.json').read().strip().splitlines()[-1])
print('$wl', 'value %.4g'%d['value'], 'ms %.4g'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'e2e ms %.4g'%d['e2e']['ms_per_step'])
PY
done
