#!/bin/bash
mkdir -p gpurun_out
R=${1:-r1y}
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$R.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$R.log; tail -3 gpurun_out/pytest_gpu_$R.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$R.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke_$R.log; tail -2 gpurun_out/smoke_$R.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -3 gpurun_out/bench_$R.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$R.json').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'], 3), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'roofline', round(d['roofline']['frac'], 3), 'loss ms', round(d['loss']['ms'], 4),
      'nms ms', round(d['nms']['ms_per_batch'], 4), d['roofline']['step_breakdown_ms'])
PY
timeout 600 python bench.py --preset vgg512 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vgg512_$R.json 2> gpurun_out/bench_vgg512_$R.err; cut -c1-200 gpurun_out/bench_vgg512_$R.json
timeout 300 python tools/quick_bench.py vgg300 64 > gpurun_out/qb_$R.log 2>&1; cp gpurun_out/quick_bench_vgg300_64.json gpurun_out/qb_$R.json; head -1 gpurun_out/qb_$R.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv python tools/ncu_target.py 64 2 > gpurun_out/ncu_launches_$R.log 2>&1; tail -1 gpurun_out/ncu_launches_$R.log
