#!/bin/bash
# round evidence: smoke, reference arm, bench, ncu launch list + full capture of the conv kernels
mkdir -p gpurun_out
R=${1:-r1}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$R.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke_$R.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
tail -2 gpurun_out/smoke_$R.log; cut -c1-300 gpurun_out/bench_ref_$R.json; cut -c1-1200 gpurun_out/bench_$R.json; tail -3 gpurun_out/bench_$R.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv python tools/ncu_target.py 64 2 > gpurun_out/ncu_launches_$R.log 2>&1
tail -1 gpurun_out/ncu_launches_$R.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 70 -c 12 -o gpurun_out/prof_conv_tc_$R python tools/ncu_target.py 64 2 > gpurun_out/ncu_full_$R.log 2>&1
tail -1 gpurun_out/ncu_full_$R.log
timeout 600 ncu --set full --clock-control none -k regex:"multibox_loss|decode_nms" -c 2 -o gpurun_out/prof_loss_$R python tools/ncu_target.py 64 1 > gpurun_out/ncu_loss_$R.log 2>&1
ls -la gpurun_out | grep $R
python bench.py --preset vgg512 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vgg512_$R.json 2> gpurun_out/bench_vgg512_$R.err
cut -c1-700 gpurun_out/bench_vgg512_$R.json; tail -3 gpurun_out/bench_vgg512_$R.err
