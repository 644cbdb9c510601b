#!/bin/bash
mkdir -p gpurun_out
R=${1:-r2c}
LOG=gpurun_out/pytest_gpu_$R.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 900 python -m pytest "$@" -q --timeout 600 -p no:cacheprovider >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run tests/test_gpu_conv.py -m gpu -x
grep -E "^===|^exit|passed|failed|^FAILED|PARITY" $LOG | cut -c1-300 | head -20
timeout 600 python tools/layer_bench.py vgg300 64 split > gpurun_out/layer_bench_$R.txt 2>&1; cat gpurun_out/layer_bench_$R.txt | cut -c1-200
echo "=== quick bench"
timeout 600 python tools/quick_bench.py vgg300 64 2>&1 | head -1 | cut -c1-700
