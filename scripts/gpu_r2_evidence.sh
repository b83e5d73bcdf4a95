#!/bin/bash
# round-2 evidence: measured FP32 rates (scripts/micro/fp32_rate), ncu --set full of the int preselection kernels, the
# exact-route kernels and the push kernel; summaries are made from the reports back in the build container
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/fp32_rate scripts/micro/fp32_rate.cu 2> gpurun_out/fp32_build.log
./scripts/micro/fp32_rate > gpurun_out/fp32_rate.txt 2>&1; echo "fp32_rate exit $?"; cat gpurun_out/fp32_rate.txt
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm --format=csv,noheader >> gpurun_out/fp32_rate.txt
for spec in "presel_int_select:gmm-presel-int" "presel_int_score:gmm-presel-int" "gmm_split_features:gmm" "gmm_tensor_kernel:gmm" "gmm_refine:gmm"; do
  k=${spec%%:*}; wl=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/r2_prof_$k \
     python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
