#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/dcp.py <<'PY'
import numpy as np, time
from rasr_b200 import flow, synth
x, offs = synth.corpus(125, n_samples=160240)
x = x.copy()
for u in range(0, 125, 3):
    a = int(offs[u]) + 40000
    x[a:a + 900] = x[a]
fe = flow.FrontEnd()
fe.process_dc(x, offs)
t = time.perf_counter(); r = fe.process_dc(x, offs, timestamps=False); t1 = time.perf_counter() - t
t = time.perf_counter(); r2 = fe.process(x, offs, timestamps=False); t2 = time.perf_counter() - t
print("process_dc %.2f ms, process %.2f ms, frames %d vs %d, runs %d" % (t1 * 1e3, t2 * 1e3, r["frame_offsets"][-1], r2["frame_offsets"][-1], len(r["runs"]["utt"])))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dc_|mfcc_" --csv --log-file gpurun_out/launches_dc.csv env PYTHONPATH=$PWD python /tmp/dcp.py > gpurun_out/dc_prof.log 2>&1
tail -2 gpurun_out/dc_prof.log
grep -E "dc_|mfcc_" gpurun_out/launches_dc.csv | awk -F'","' '{print $5, $NF}' | tail -12
grep -E "dc_" gpurun_out/launches_dc.csv | awk -F'","' '{print $5, $NF}'
timeout 600 python -m pytest tests/test_gpu_frontend_dc.py -q -m gpu --tb=short -x 2>&1 | tail -3
