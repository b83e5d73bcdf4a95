#!/bin/bash
# N-GPU launch exactly as the driver does it (torchrun, one rank per GPU), both arms
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"
tail -n 1 gpurun_out/bench_n$N.json | cut -c1-600; tail -n 3 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref N=$N exit $?"
tail -n 1 gpurun_out/bench_ref_n$N.json | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --workload pipeline > gpurun_out/bench_pipe_n$N.json 2> gpurun_out/bench_pipe_n$N.err; echo "pipeline N=$N exit $?"
tail -n 1 gpurun_out/bench_pipe_n$N.json | cut -c1-400
