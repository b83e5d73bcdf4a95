#!/bin/bash
# all GPU tests + every bench workload (profiles for DESIGN.md)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --tb=short > gpurun_out/test_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -n 5 gpurun_out/test_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
for wl in gmm gmm-int gmm-tensor gmm-diag gmm-presel gmm-presel-int frontend pipeline pipeline-nn pipeline-search nn; do
  extra=""; [ $wl = nn ] && extra="--frames 75776 --steps 10"
  timeout 600 python bench.py --workload $wl $extra > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl exit $?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$wl.json').read().strip().splitlines()[-1])
print('$wl', 'value %.4g'%d['value'], 'ms %.4g'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'roof %.4g %s frac %.4g'%(d['roofline']['achieved'], d['roofline']['unit'], d['roofline']['frac']), d.get('cpu_baseline',{}).get('value'), {k:v['value'] for k,v in d.get('variants',{}).items()})
PY
done
