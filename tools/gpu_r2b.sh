#!/bin/bash
mkdir -p gpurun_out
R=${1:-r2b}
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$R.json'))
print({k:d[k] for k in ('value','ms_per_step','clocks','gpu_launches')}, d['e2e'])
print(d['roofline']['step_breakdown_ms'])
PY
timeout 600 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/ncu_step_$R.csv python tools/ncu_step.py vgg300 64 train > gpurun_out/ncu_step_$R.log 2>&1
tail -2 gpurun_out/ncu_step_$R.log; wc -l gpurun_out/ncu_step_$R.csv
SSDB_CONV=tf32 timeout 300 python tools/quick_bench.py vgg300 64 2>&1 | head -1 | cut -c1-400
