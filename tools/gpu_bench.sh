#!/bin/bash
# bench + ncu evidence for one round; outputs under gpurun_out/
mkdir -p gpurun_out
R=${1:-r1}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$R.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke_$R.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
tail -3 gpurun_out/smoke_$R.log; cat gpurun_out/bench_ref_$R.json; cat gpurun_out/bench_$R.json; tail -5 gpurun_out/bench_$R.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv python tools/ncu_target.py 16 2 > gpurun_out/ncu_launches_$R.log 2>&1
tail -2 gpurun_out/ncu_launches_$R.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 60 -c 6 -o gpurun_out/prof_conv_tc_$R python tools/ncu_target.py 16 2 > gpurun_out/ncu_full_$R.log 2>&1
tail -2 gpurun_out/ncu_full_$R.log
ls -la gpurun_out | tail -12
