#!/bin/bash
# end-of-round evidence: the driver's own sequence (pytest -m gpu as one process, smoke, reference arm, bench) + ncu captures
mkdir -p gpurun_out
R=${1:-r1z}
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$R.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$R.log; tail -3 gpurun_out/pytest_gpu_$R.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$R.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke_$R.log; tail -2 gpurun_out/smoke_$R.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err; cut -c1-300 gpurun_out/bench_ref_$R.json
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -3 gpurun_out/bench_$R.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$R.json').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'], 3), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'roofline', round(d['roofline']['frac'], 3), 'loss ms', round(d['loss']['ms'], 4), 'frac', round(d['loss']['roofline']['frac'], 3),
      'nms ms', round(d['nms']['ms_per_batch'], 4), 'frac', round(d['nms']['roofline']['frac'], 3), 'cpu', d['cpu_baseline'], 'clocks', d['clocks'])
PY
timeout 600 python bench.py --preset vgg512 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vgg512_$R.json 2> gpurun_out/bench_vgg512_$R.err; cut -c1-200 gpurun_out/bench_vgg512_$R.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv python tools/ncu_target.py 64 2 > gpurun_out/ncu_launches_$R.log 2>&1; tail -1 gpurun_out/ncu_launches_$R.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_box_$R.csv python tools/ncu_target_box.py 2 > gpurun_out/ncu_launches_box_$R.log 2>&1; tail -1 gpurun_out/ncu_launches_box_$R.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"loss_rows|loss_select|loss_grad|match_best|decode_scan|decode_nms" -c 9 -o gpurun_out/prof_box_$R python tools/ncu_target_box.py 1 > gpurun_out/ncu_full_box_$R.log 2>&1; tail -1 gpurun_out/ncu_full_box_$R.log
