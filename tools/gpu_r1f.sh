#!/bin/bash
mkdir -p gpurun_out
R=${1:-r1f}
LOG=gpurun_out/pytest_gpu_$R.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 900 python -m pytest "$@" -q --timeout 600 -p no:cacheprovider >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run tests/test_gpu_box.py tests/test_gpu_loss.py -m gpu
grep -E "^===|^exit|passed|failed|Error|error|assert" $LOG | cut -c1-300 | head -40
SSDB_TRACE=1 timeout 200 python tools/ncu_target_box.py 1 2>&1 | tail -6
timeout 120 python tools/nms_diag.py 2>&1 | cut -c1-200 | head -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
tail -3 gpurun_out/bench_$R.err
python - <<PY
import json
for f in ('gpurun_out/bench_$R.json',):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'loss ms', d['loss']['ms'], 'frac', d['loss']['roofline']['frac'],
              'nms ms', d['nms']['ms_per_batch'], 'frac', d['nms']['roofline']['frac'], 'nms e2e ms', d['nms']['e2e']['ms_per_batch'])
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
