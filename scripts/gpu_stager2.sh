#!/bin/bash
# A/B on ONE box: host stager off / on, C2 and C3, two rounds each (hosts differ from box to box)
mkdir -p gpurun_out
for round in 1 2; do
for wl in gmm pipeline; do
for mode in off on; do
  if [ $mode = off ]; then export RB_NO_HOST_STAGER=1; else unset RB_NO_HOST_STAGER; fi
  timeout 300 python bench.py --workload $wl --no-cpu-baseline --steps 10 > gpurun_out/ab.json 2>/dev/null
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/ab.json') if l.startswith('{')][-1])
print("$wl stager $mode round $round: e2e %.1fM pageable %.1fM ceiling %.1fM" % (d['e2e']['value']/1e6, d['e2e']['pageable']['value']/1e6, d['e2e']['pcie_ceiling']['value']/1e6))
PY
done; done; done
