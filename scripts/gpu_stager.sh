#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --tb=short > gpurun_out/test_gpu.log 2>&1
echo "pytest -m gpu exit $?"; tail -n 4 gpurun_out/test_gpu.log
for t in 4; do
  RB_COPY_THREADS=$t timeout 300 python bench.py --workload gmm --no-cpu-baseline --steps 10 > gpurun_out/bench_stager_$t.json 2> gpurun_out/bench_stager_$t.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_stager_$t.json') if l.startswith('{')][-1])
print("threads $t", "value %.1fM e2e %.1fM pageable %.1fM" % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e']['pageable']['value']/1e6))
PY
done
RB_NO_HOST_STAGER=1 timeout 300 python bench.py --workload pipeline --no-cpu-baseline --steps 5 > gpurun_out/bench_stager_off.json 2>/dev/null
timeout 300 python bench.py --workload pipeline --no-cpu-baseline --steps 5 > gpurun_out/bench_stager_on.json 2>/dev/null
python - <<PY
import json
for n in ("off","on"):
    d=json.loads([l for l in open('gpurun_out/bench_stager_%s.json'%n) if l.startswith('{')][-1])
    print("pipeline stager", n, "e2e %.1fM pageable %.1fM" % (d['e2e']['value']/1e6, d['e2e']['pageable']['value']/1e6))
PY
nproc
