#!/bin/bash
# compute-sanitizer memcheck over the quick parity tests of the kernels added last (out-of-bounds / misaligned accesses)
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitize_$name.log python -m pytest "$@" -q -m gpu -x --tb=line > gpurun_out/sanitize_$name.out 2>&1
  echo "$name exit $?"; tail -n 2 gpurun_out/sanitize_$name.out; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/sanitize_$name.log; tail -n 2 gpurun_out/sanitize_$name.log
}
run presel tests/test_gpu_gmm_presel.py -k "small_models or other_clustering or rejects"
run dc tests/test_gpu_frontend_dc.py -k "batch_with_dc or other_parameters or float_input"
run search tests/test_gpu_search.py -k "kernel_variants or bit_exact"
run postproc tests/test_gpu_postproc.py
