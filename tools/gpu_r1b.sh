#!/bin/bash
# second GPU pass of round 1: new streaming loss / decode+NMS kernels -> tests, bench (v2 and v1), ncu
mkdir -p gpurun_out
R=${1:-r1b}
LOG=gpurun_out/pytest_gpu_$R.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 900 python -m pytest "$@" -q --timeout 600 -p no:cacheprovider >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run tests/test_gpu_box.py tests/test_gpu_loss.py -m gpu
run tests -x -m gpu --deselect tests/test_gpu_box.py --deselect tests/test_gpu_loss.py
grep -E "^===|^exit|passed|failed|Error|error|assert" $LOG | cut -c1-300 | head -60
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
cut -c1-3000 gpurun_out/bench_$R.json; tail -3 gpurun_out/bench_$R.err
SSDB_LOSS=v1 SSDB_NMS=v1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${R}_v1.json 2> gpurun_out/bench_${R}_v1.err
python - <<PY
import json
for f in ('gpurun_out/bench_$R.json', 'gpurun_out/bench_${R}_v1.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'loss ms', d['loss']['ms'], 'frac', d['loss']['roofline']['frac'],
              'nms ms', d['nms']['ms_per_batch'], 'frac', d['nms']['roofline']['frac'])
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv python tools/ncu_target.py 64 2 > gpurun_out/ncu_launches_$R.log 2>&1
tail -1 gpurun_out/ncu_launches_$R.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_box_$R.csv python tools/ncu_target_box.py 2 > gpurun_out/ncu_launches_box_$R.log 2>&1
tail -2 gpurun_out/ncu_launches_box_$R.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"loss_rows|loss_select|loss_grad|match_best|decode_scan|decode_nms" -c 12 -o gpurun_out/prof_box_$R python tools/ncu_target_box.py 1 > gpurun_out/ncu_full_box_$R.log 2>&1
tail -1 gpurun_out/ncu_full_box_$R.log
ls -la gpurun_out | grep $R
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$R.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke_$R.log; tail -2 gpurun_out/smoke_$R.log
