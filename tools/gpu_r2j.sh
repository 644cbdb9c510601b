#!/bin/bash
mkdir -p gpurun_out
R=${1:-r2j}
LOG=gpurun_out/pytest_gpu_$R.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 900 python -m pytest "$@" -q --timeout 600 -p no:cacheprovider >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run tests/test_gpu_surface.py -m gpu
run tests/test_gpu_conv.py -m gpu -x
grep -E "^===|^exit|passed|failed|^FAILED|^ERROR|^E  " $LOG | cut -c1-300 | head -30
timeout 300 python tools/layer_bench.py vgg300 64 split > gpurun_out/layer_bench_$R.txt 2>&1; tail -1 gpurun_out/layer_bench_$R.txt
timeout 300 python tools/quick_bench.py vgg300 64 2>&1 | head -1 | cut -c1-600
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0,'ssd-tensorflow_b200')
import ssdvgg, synth, ssdutils
for frozen in (False, True):
    sess=ssdvgg.Session(); m=ssdvgg.SSDVGG(sess, ssdutils.get_preset_by_name('vgg300')); m.build_from_vgg(None,20); m._frozen=frozen
    for B in (1, 8, 64):
        x=synth.images(0,B,300)
        import torch
        x=torch.from_numpy(x).pin_memory().numpy()
        for _ in range(3): m.detect(x,0.01,{},200,rows=True)
        t0=time.perf_counter()
        for _ in range(20): m.detect(x,0.01,{},200,rows=True)
        print('frozen' if frozen else 'eager ', 'B=%d detect %.3f ms/batch'%(B,(time.perf_counter()-t0)/20*1e3))
    sess.close()
PY
