#!/bin/bash
mkdir -p gpurun_out
R=${1:-r1j}
LOG=gpurun_out/pytest_gpu_$R.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 900 python -m pytest "$@" -q --timeout 600 -p no:cacheprovider >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run tests/test_gpu_box.py tests/test_gpu_loss.py -m gpu
grep -E "^===|^exit|passed|failed|Error|error|assert" $LOG | cut -c1-300 | head -40
SSDB_TRACE=1 timeout 200 python tools/ncu_target_box.py 1 2>&1 | grep -v "^  walk" | tail -6
timeout 120 python tools/nms_diag.py 2>&1 | cut -c1-200 | head -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_box_$R.csv python tools/ncu_target_box.py 2 > gpurun_out/ncu_launches_box_$R.log 2>&1
python tools/ncu_summary.py launches gpurun_out/launches_box_$R.csv gpurun_out/launches_box_$R.txt | head -14
