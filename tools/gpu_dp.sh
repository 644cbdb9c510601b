#!/bin/bash
# 2-GPU data-parallel bench (run with: gpurun --gpus 2)
mkdir -p gpurun_out
R=${1:-r1}
N=${2:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${R}_n$N.json 2> gpurun_out/bench_${R}_n$N.err
cat gpurun_out/bench_${R}_n$N.json | cut -c1-900; tail -5 gpurun_out/bench_${R}_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 tools/dp_parity.py > gpurun_out/dp_parity_${R}_n$N.log 2>&1
tail -5 gpurun_out/dp_parity_${R}_n$N.log
