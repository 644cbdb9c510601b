#!/bin/bash
# round-2 bring-up of the split-operand kernels
mkdir -p gpurun_out
R=${1:-r2a}
timeout 300 python tools/split_bringup.py > gpurun_out/bringup_$R.log 2>&1; echo "bringup exit $?"
cat gpurun_out/bringup_$R.log | cut -c1-260 | tail -40
LOG=gpurun_out/pytest_gpu_$R.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 900 python -m pytest "$@" -q --timeout 600 -p no:cacheprovider >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run tests/test_gpu_conv.py -m gpu
run tests/test_gpu_net.py -m gpu -s
grep -E "^===|^exit|passed|failed|^FAILED|PARITY" $LOG | cut -c1-600 | head -80
echo "=== quick bench"
timeout 600 python tools/quick_bench.py vgg300 64 2>&1 | tail -34 | cut -c1-200
