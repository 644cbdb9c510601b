#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_n$N.json 2> gpurun_out/bench_r1_n$N.err
cut -c1-900 gpurun_out/bench_r1_n$N.json; tail -4 gpurun_out/bench_r1_n$N.err | cut -c1-300
