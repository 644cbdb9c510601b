#!/bin/bash
# A/B of the two-CTAs-per-SM mode of conv_tc_kernel (SSDB_TC_DUAL=1): parity tests, then per-layer timings
mkdir -p gpurun_out
export SSDB_TC_DUAL=1
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_net.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
timeout 300 python tools/quick_bench.py vgg300 64 > gpurun_out/qb_dual1.log 2>&1; cp gpurun_out/quick_bench_vgg300_64.json gpurun_out/qb_dual1.json
export SSDB_TC_DUAL=0
timeout 300 python tools/quick_bench.py vgg300 64 > gpurun_out/qb_dual0.log 2>&1; cp gpurun_out/quick_bench_vgg300_64.json gpurun_out/qb_dual0.json
python - <<PY
import json
a = json.load(open('gpurun_out/qb_dual0.json')); b = json.load(open('gpurun_out/qb_dual1.json'))
print('train ms', round(a['train_ms'], 3), '->', round(b['train_ms'], 3), ' fwd ms', round(a['fwd_ms'], 3), '->', round(b['fwd_ms'], 3))
pa = {l: m for l, m, _ in a['profile']}
for l, m, _ in b['profile']:
    if abs(m - pa.get(l, m)) > 0.03: print('%-28s %7.3f -> %7.3f' % (l, pa[l], m))
PY
