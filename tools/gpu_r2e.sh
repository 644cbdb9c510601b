#!/bin/bash
mkdir -p gpurun_out
R=${1:-r2e}
LOG=gpurun_out/pytest_gpu_$R.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 900 python -m pytest "$@" -q --timeout 600 -p no:cacheprovider >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run tests/test_gpu_conv.py -m gpu
grep -E "^===|^exit|passed|failed|^FAILED|^ERROR" $LOG | cut -c1-300 | head -20
echo "=== wgrad r2 on / off"
timeout 300 python tools/layer_bench.py vgg300 64 split conv2 conv3_2 conv4_2 conv4_1 head0 2>&1 | cut -c1-200
SSDB_WG_R2=0 timeout 300 python tools/layer_bench.py vgg300 64 split conv2 conv3_2 conv4_2 head0 2>&1 | cut -c1-200
echo "=== rw mtu2"
SSDB_RW_MTU2=1 timeout 300 python tools/layer_bench.py vgg300 64 split conv2 head0 2>&1 | cut -c1-200
echo "=== bench"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -3 gpurun_out/bench_$R.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$R.json'))
for k in ('value','ms_per_step','clocks','gpu_launches','e2e','e2e_gt_feed','cpu_baseline'): print(k, d.get(k))
print('roofline', {k:v for k,v in d['roofline'].items() if k not in ('step_breakdown_ms',)}); print(d['roofline']['step_breakdown_ms'])
for k in ('tf32_mode','vgg512','forward_only'):
    if k in d: print(k, json.dumps(d[k])[:900])
print('loss', json.dumps(d['loss'])[:600]); print('nms', json.dumps(d['nms'])[:900])
PY
