#!/bin/bash
mkdir -p gpurun_out
R=${1:-r2h}
LOG=gpurun_out/pytest_gpu_$R.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 900 python -m pytest "$@" -q --timeout 600 -p no:cacheprovider >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run tests/test_gpu_conv.py tests/test_gpu_net.py -m gpu
grep -E "^===|^exit|passed|failed|^FAILED|^ERROR" $LOG | cut -c1-300 | head -20
echo "=== layer bench, 8 epilogue warps"
timeout 300 python tools/layer_bench.py vgg300 64 split > gpurun_out/layer_bench_$R.txt 2>&1; cut -c1-170 gpurun_out/layer_bench_$R.txt
echo "=== bench"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -3 gpurun_out/bench_$R.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$R.json'))
for k in ('value','ms_per_step','clocks','gpu_launches'): print(k, d.get(k))
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'gt', d['e2e_gt_feed']['value'], d['e2e_gt_feed']['ms_per_step'])
print('roofline frac', d['roofline']['frac'], d['roofline']['step_breakdown_ms'])
print('vgg512', d['vgg512']['value'], d['vgg512']['e2e']['value'], d['vgg512']['roofline']['frac'])
PY
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_ncu_$R.log 2>&1; wc -l gpurun_out/launches_$R.csv
echo "=== ncu full, labelled layers (report kept on the box, raw page exported)"
timeout 900 ncu --set full --clock-control none -k regex:conv_tc -o /tmp/prof_layers -f python tools/layer_bench.py vgg300 64 split conv4_2 conv1_2 conv2_2 --once > gpurun_out/ncu_layers_$R.log 2>&1
ncu -i /tmp/prof_layers.ncu-rep --page raw --csv > gpurun_out/prof_layers_raw_$R.csv 2>/dev/null; wc -c gpurun_out/prof_layers_raw_$R.csv; du -sh gpurun_out
