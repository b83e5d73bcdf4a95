#!/bin/bash
# Nn GEMM: parity tests, the C4 bench line and a launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nn.py tests/test_gpu_pipeline.py -q -x --tb=short > gpurun_out/test_nn.log 2>&1
echo "nn tests exit $?"; tail -n 4 gpurun_out/test_nn.log
timeout 300 python bench.py --workload nn --no-cpu-baseline > gpurun_out/bench_nn.json 2> gpurun_out/bench_nn.err
echo "bench exit $?"; cut -c1-700 gpurun_out/bench_nn.json; tail -n 3 gpurun_out/bench_nn.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_nn.csv \
    python bench.py --workload nn --no-cpu-baseline --steps 3 --warmup 3 --frames 18944 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/launches_nn.csv') if not l.startswith('=='))]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:10]:
    print(r[ki][:90].ljust(90), r[vi])
PY
