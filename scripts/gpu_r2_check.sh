#!/bin/bash
# round 2: full GPU suite + smoke + both bench arms, then compute-sanitizer on the kernels added since r1_sanitizer.md
mkdir -p gpurun_out
bash scripts/gpu_full.sh
run() {
  tool=$1; name=$2; shift 2
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/${tool}_$name.log python -m pytest "$@" -q -m gpu -x --tb=line > gpurun_out/${tool}_$name.out 2>&1
  echo "$tool $name exit $?" | tee -a gpurun_out/summary.txt; tail -n 1 gpurun_out/${tool}_$name.out; tail -n 1 gpurun_out/${tool}_$name.log
}
run memcheck exact tests/test_gpu_gmm_exact.py -k "not full_c2 and not screening_error"
run memcheck preselint tests/test_gpu_zz_gmm_presel_int.py -k "not 100k and not full"
run memcheck preseltie tests/test_gpu_gmm_presel.py -k "tie"
run memcheck comm tests/test_gpu_multi.py -k "single_rank"
run racecheck exact tests/test_gpu_gmm_exact.py -k "ragged or uniform"
run racecheck preselint tests/test_gpu_zz_gmm_presel_int.py -k "small or tie"
