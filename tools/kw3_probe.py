"""Bring-up: which shared-memory addressing does the tensor core accept for row windows into a wider TMA box?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ssd-tensorflow_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import ssdb
from gpu_util import conv_case, rel_err, run_fprop
for boxw, bo in ((10, 1), (10, 0), (16, 1), (16, 0)):
    os.environ['SSDB_KW3_BOXW'] = str(boxw); os.environ['SSDB_KW3_BO'] = str(bo)
    line = 'boxw %d base_offset %d:' % (boxw, bo)
    for (B, H, Cin, Cout) in ((2, 64, 64, 64), (2, 150, 64, 128)):
        x, w, b, pad, Ho = conv_case(B, H, Cin, Cout, 3, 1, 1, 'SAME', seed=1)
        ys = run_fprop(ssdb.CONV_SIMT, x, w, b, 3, 1, 1, pad, Ho)
        os.environ['SSDB_KW3'] = '1'
        yt = run_fprop(ssdb.CONV_TC, x, w, b, 3, 1, 1, pad, Ho)
        line += '  %dx%d %d->%d err %.4f' % (H, H, Cin, Cout, rel_err(yt, ys))
    print(line, flush=True)
os.environ['SSDB_KW3'] = '0'
x, w, b, pad, Ho = conv_case(2, 64, 64, 64, 3, 1, 1, 'SAME', seed=1)
print('plain mode err %.5f' % rel_err(run_fprop(ssdb.CONV_TC, x, w, b, 3, 1, 1, pad, Ho), run_fprop(ssdb.CONV_SIMT, x, w, b, 3, 1, 1, pad, Ho)))
