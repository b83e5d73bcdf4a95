#!/bin/bash
# gpurun --gpus N: the driver's scaling line at N ranks (default workloads + the score exchange), after the 2-rank tests
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x --tb=short > gpurun_out/test_multi.log 2>&1
echo "multi tests exit $?" | tee gpurun_out/summary_multi.txt
tail -n 5 gpurun_out/test_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N > gpurun_out/bench_default_n$N.json 2> gpurun_out/bench_default_n$N.err
echo "bench n$N exit $?" | tee -a gpurun_out/summary_multi.txt
tail -c 3000 gpurun_out/bench_default_n$N.json; tail -n 5 gpurun_out/bench_default_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference_n$N.json 2> gpurun_out/bench_reference_n$N.err
echo "bench ref n$N exit $?" | tee -a gpurun_out/summary_multi.txt
