#!/bin/bash
# third GPU pass: contention-free select kernels -> box/loss tests, bench, ncu of the box kernels
mkdir -p gpurun_out
R=${1:-r1c}
LOG=gpurun_out/pytest_gpu_$R.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 900 python -m pytest "$@" -q --timeout 600 -p no:cacheprovider >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run tests/test_gpu_box.py tests/test_gpu_loss.py -m gpu
run tests/test_gpu_surface.py tests/test_gpu_net.py -x -m gpu
grep -E "^===|^exit|passed|failed|Error|error|assert" $LOG | cut -c1-300 | head -60
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
tail -3 gpurun_out/bench_$R.err
python - <<PY
import json
for f in ('gpurun_out/bench_$R.json',):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'loss ms', d['loss']['ms'], 'frac', d['loss']['roofline']['frac'],
              'nms ms', d['nms']['ms_per_batch'], 'frac', d['nms']['roofline']['frac'], 'nms e2e ms', d['nms']['e2e']['ms_per_batch'], 'nms cpu', d['nms']['cpu_baseline'])
        print(d['roofline']['step_breakdown_ms'])
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_box_$R.csv python tools/ncu_target_box.py 2 > gpurun_out/ncu_launches_box_$R.log 2>&1
tail -2 gpurun_out/ncu_launches_box_$R.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"loss_rows|loss_select|loss_grad|match_best|decode_scan|decode_nms" -c 12 -o gpurun_out/prof_box_$R python tools/ncu_target_box.py 1 > gpurun_out/ncu_full_box_$R.log 2>&1
tail -1 gpurun_out/ncu_full_box_$R.log
