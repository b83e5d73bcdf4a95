#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gmm_exact.py tests/test_gpu_gmm.py -q -x --tb=short > gpurun_out/test_exact_diag.log 2>&1
echo "tests exit $?"; tail -n 30 gpurun_out/test_exact_diag.log
for e in 0 1; do
RB_GMM_EXACT=$e timeout 300 python bench.py --workload gmm-diag --no-cpu-baseline --steps 10 > gpurun_out/bench_diag_$e.json 2> gpurun_out/bench_diag_$e.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_diag_$e.json') if l.startswith('{')][-1])
print("RB_GMM_EXACT=$e gmm-diag value %.1fM ms %.3f e2e %.1fM" % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_exact_diag.csv \
    python bench.py --workload gmm-diag --no-cpu-baseline --steps 3 --warmup 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/launches_exact_diag.csv') if not l.startswith('=='))]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:7]:
    print(r[ki][:80].ljust(80), r[vi])
PY
