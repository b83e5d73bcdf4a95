#!/bin/bash
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitize_$name.log python -m pytest "$@" -q -m gpu -x --tb=line > gpurun_out/sanitize_$name.out 2>&1
  echo "$name exit $?"; tail -n 1 gpurun_out/sanitize_$name.out; tail -n 1 gpurun_out/sanitize_$name.log
}
run gmm tests/test_gpu_gmm.py -k "not full_size and not 100k"
run gmmint tests/test_gpu_gmm_int.py -k "dimensions or ragged"
run frontend tests/test_gpu_frontend.py -k "c1_utterance or chunk or s16 or batch"
run nn tests/test_gpu_nn.py -k "unit_test or two_layer or f32_path or bf16_path"
