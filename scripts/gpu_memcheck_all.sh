#!/bin/bash
# compute-sanitizer memcheck over the whole GPU suite (minus the full-size cases), one pytest process per file so that a
# slow file cannot starve the others; summary lines -> gpurun_out/memcheck_all.txt
mkdir -p gpurun_out
: > gpurun_out/memcheck_all.txt
for f in tests/test_gpu_*.py; do
  name=$(basename $f .py)
  timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_$name.log \
      python -m pytest $f -q -m gpu -x --tb=line -k "not full and not 100k and not full_size and not c2_batch and not reference_host" > gpurun_out/memcheck_$name.out 2>&1
  rc=$?
  echo "$name rc=$rc | $(tail -n 1 gpurun_out/memcheck_$name.out) | $(grep -h 'ERROR SUMMARY' gpurun_out/memcheck_$name.log | tail -n 1)" | tee -a gpurun_out/memcheck_all.txt
done
