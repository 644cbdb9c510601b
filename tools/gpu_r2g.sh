#!/bin/bash
mkdir -p gpurun_out
R=${1:-r2g}
bash tools/gpu_tests.sh $R 2>&1 | tail -30
echo "=== bench"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -3 gpurun_out/bench_$R.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$R.json'))
for k in ('value','ms_per_step','clocks','gpu_launches'): print(k, d.get(k))
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'gt', d['e2e_gt_feed']['value'], d['e2e_gt_feed']['ms_per_step'])
print('roofline frac', d['roofline']['frac'], d['roofline']['step_breakdown_ms'])
print('vgg512', d['vgg512']['value'], d['vgg512']['e2e']['value'], d['vgg512']['roofline']['frac'])
print('tf32', d['tf32_mode']['value']); print('fwd', json.dumps(d['forward_only'])[:600])
print('nms e2e', d['nms']['e2e']['ms_per_batch'])
PY
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_ncu_$R.log 2>&1; wc -l gpurun_out/launches_$R.csv
echo "=== ncu full, labelled layers"
timeout 900 ncu --set full --clock-control none -k regex:conv_tc -o gpurun_out/prof_layers_$R -f python tools/layer_bench.py vgg300 64 split conv4_2 conv1_2 conv2_2 --once > gpurun_out/ncu_layers_$R.log 2>&1; tail -4 gpurun_out/ncu_layers_$R.log; ls -la gpurun_out/prof_layers_$R.ncu-rep
