#!/bin/bash
# gpurun --gpus 2: the score exchange (rb_comm_*) tests and its bench lines
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x --tb=short > gpurun_out/test_multi.log 2>&1
echo "multi tests exit $?" | tee gpurun_out/summary_multi.txt
tail -n 30 gpurun_out/test_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 --workload gmm --gather > gpurun_out/bench_gather_n2.json 2> gpurun_out/bench_gather_n2.err
echo "bench gather exit $?" | tee -a gpurun_out/summary_multi.txt
cat gpurun_out/bench_gather_n2.json; tail -n 5 gpurun_out/bench_gather_n2.err
