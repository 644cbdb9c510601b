#!/bin/bash
mkdir -p gpurun_out
R=${1:-r2k}
bash tools/gpu_tests.sh $R 2>&1 | grep -E "^===|^exit|passed|failed|^FAILED|^ERROR" | head -30
echo "=== dual stream on / off"
timeout 300 python tools/quick_bench.py vgg300 64 2>&1 | head -1 | cut -c1-330
SSDB_DUAL_STREAM=0 timeout 300 python tools/quick_bench.py vgg300 64 2>&1 | head -1 | cut -c1-330
timeout 300 python tools/quick_bench.py vgg512 32 2>&1 | head -1 | cut -c1-330
SSDB_DUAL_STREAM=0 timeout 300 python tools/quick_bench.py vgg512 32 2>&1 | head -1 | cut -c1-330
python -c "
import __graft_entry__ as g
g.smoke()
"
