#!/bin/bash
# N-GPU data-parallel bench only (run with: gpurun --gpus N)
mkdir -p gpurun_out
R=${1:-r2}
N=${2:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${R}_n$N.json 2> gpurun_out/bench_${R}_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${R}_n$N.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','clocks','gpu_launches')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'gt', d['e2e_gt_feed']['value'], d['e2e_gt_feed']['ms_per_step'])
PY
tail -2 gpurun_out/bench_${R}_n$N.err
