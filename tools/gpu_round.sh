#!/bin/bash
# tests (grouped per process) + quick bench + bench.py ; outputs under gpurun_out/
mkdir -p gpurun_out
LOG=gpurun_out/pytest_gpu.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 900 python -m pytest "$@" -q -s --timeout 600 -p no:cacheprovider >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run tests/test_gpu_box.py tests/test_gpu_loss.py -m gpu
run tests/test_gpu_conv.py -m gpu
run tests/test_gpu_net.py -m gpu
run tests/test_gpu_surface.py -m gpu
grep -E "^===|^exit|passed|failed|Error|error|PARITY" $LOG | cut -c1-420 | head -60
echo "=== quick bench"
timeout 600 python tools/quick_bench.py vgg300 64 2>&1 | tail -32 | cut -c1-200
R=${1:-r1}
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
cat gpurun_out/bench_$R.json; tail -5 gpurun_out/bench_$R.err
