#!/bin/bash
# GPU bring-up: every gpu test group in its own process (a trapped kernel poisons the context)
mkdir -p gpurun_out
LOG=gpurun_out/pytest_gpu.log
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 900 python -m pytest "$@" -q -s --timeout 600 -p no:cacheprovider >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run tests/test_gpu_box.py -m gpu
run tests/test_gpu_loss.py -m gpu
run tests/test_gpu_conv.py -m gpu -k simt
run tests/test_gpu_conv.py -m gpu -k tcgen05
run tests/test_gpu_net.py -m gpu -k simt
run tests/test_gpu_net.py -m gpu -k auto
grep -E "^===|^exit|passed|failed|Error|error|PARITY" $LOG | cut -c1-300 | head -80
echo "=== quick bench" >> $LOG
timeout 600 python tools/quick_bench.py vgg300 32 >> $LOG 2>&1
timeout 600 python tools/quick_bench.py vgg300 64 >> $LOG 2>&1
tail -70 $LOG | cut -c1-200
