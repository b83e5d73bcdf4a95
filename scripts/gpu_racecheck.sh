#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over small cases of the kernels with hand-rolled synchronisation
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 --log-file gpurun_out/race_$name.log python -m pytest "$@" -q -m gpu -x --tb=line > gpurun_out/race_$name.out 2>&1
  echo "$name exit $?"; tail -n 1 gpurun_out/race_$name.out; grep -c "Race reported\|hazard" gpurun_out/race_$name.log; tail -n 3 gpurun_out/race_$name.log
}
run search tests/test_gpu_search.py -k "linear_search_bit_exact or ties_resolve"
run presel tests/test_gpu_gmm_presel.py -k "small_models"
run dc tests/test_gpu_frontend_dc.py -k "other_parameters"
