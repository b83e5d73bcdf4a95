#!/bin/bash
# evidence for the kernels added late in round 2: launch list of the C5 step (continuous + single-word search), ncu
# --set full of linear_search_reg_kernel<8, true> and of gmm_simd_kernel (DP4A path: scores + densities)
mkdir -p gpurun_out
cat > gpurun_out/simd_probe.py <<'PY'
import sys; sys.path.insert(0, ".")
import numpy as np
from rasr_b200 import mm, synth
sc = mm.GmmScorer(mm.MixtureSet.from_dict(synth.mixture_set()), "SIMD-diagonal-maximum")
f = synth.features(100000, 39)
for _ in range(4):
    sc.score(f, want_density=True)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_search_launches.csv \
    python bench.py --workload pipeline-search --steps 2 --warmup 1 > gpurun_out/ncu_search_list.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_search_reg -s 9 -c 1 -f -o gpurun_out/r2_prof_search_single \
    python bench.py --workload pipeline-search --steps 2 --warmup 1 > gpurun_out/ncu_search_single.log 2>&1
echo "ncu search single exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gmm_simd_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_gmm_simd \
    python gpurun_out/simd_probe.py > gpurun_out/ncu_gmm_simd.log 2>&1
echo "ncu gmm_simd exit $?"
ls -la gpurun_out/*.ncu-rep
