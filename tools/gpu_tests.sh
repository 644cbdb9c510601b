#!/bin/bash
# the whole -m gpu suite, grouped per process; log under gpurun_out/
mkdir -p gpurun_out
R=${1:-r2}
LOG=gpurun_out/pytest_gpu_$R.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 900 python -m pytest "$@" -q --timeout 600 -p no:cacheprovider >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run tests/test_gpu_box.py tests/test_gpu_loss.py -m gpu
run tests/test_gpu_conv.py -m gpu
run tests/test_gpu_net.py -m gpu -s
run tests/test_gpu_surface.py -m gpu
ls tests/test_gpu_grad.py >/dev/null 2>&1 && run tests/test_gpu_grad.py -m gpu -s
grep -E "^===|^exit|passed|failed|^FAILED|^ERROR|PARITY |GRAD " $LOG | cut -c1-500 | head -60
