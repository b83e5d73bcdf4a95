#!/bin/bash
# int preselection scorer: parity tests, bench line and launch list (profiles/r1_bench_gmm-presel-int.json,
# profiles/r1_gmm_presel_int_launches.csv)
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_zz_gmm_presel_int.py tests/test_gpu_gmm_presel.py -q -m gpu -p no:cacheprovider > gpurun_out/presel_int_tests.log 2>&1; tail -n 3 gpurun_out/presel_int_tests.log
timeout 60 python bench.py --workload gmm-presel-int --steps 10 --warmup 3 > gpurun_out/bench_gmm-presel-int.json 2> gpurun_out/bench_gmm-presel-int.err
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_presel_int.csv python bench.py --workload gmm-presel-int --steps 2 --warmup 3 > gpurun_out/presel_int_ncu.log 2>&1
cut -c1-300 gpurun_out/bench_gmm-presel-int.json
# still to run: memcheck / racecheck of the preselection kernels on the small cases
# timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_zz_gmm_presel_int.py -q -m gpu -k "other_clustering or constant or golden or float_variant or small_models" -p no:cacheprovider > gpurun_out/sanitizer_presel_int_memcheck.log 2>&1
# timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_zz_gmm_presel_int.py -q -m gpu -k "constant or golden or float_variant" -p no:cacheprovider > gpurun_out/sanitizer_presel_int_racecheck.log 2>&1
