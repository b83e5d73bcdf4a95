#!/bin/bash
# N-GPU launch (torchrun, one rank per GPU) of the C4 (Nn) and C5 (audio -> scores -> search) workloads
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
for wl in nn pipeline-search; do
  extra=""; [ $wl = nn ] && extra="--frames 75776"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --workload $wl $extra > gpurun_out/bench_${wl}_n$N.json 2> gpurun_out/bench_${wl}_n$N.err; echo "$wl N=$N exit $?"
  tail -n 1 gpurun_out/bench_${wl}_n$N.json | cut -c1-700; tail -n 2 gpurun_out/bench_${wl}_n$N.err
done
