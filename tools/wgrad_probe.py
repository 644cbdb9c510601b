"""Bring-up experiment: where does the tcgen05 wgrad kernel spend its time?  (SSDB_WG_DEBUG switches)"""
import os, sys, ctypes
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ssd-tensorflow_b200'))
import ssdb
L = ssdb.lib()
def P(t): return ctypes.c_void_p(t.data_ptr())
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
shapes = [(64, 38, 512, 512, 3), (64, 300, 64, 64, 3), (64, 75, 256, 256, 3), (64, 19, 512, 1024, 3)]
for (B, H, Cin, Cout, k) in shapes:
    x = torch.randn((B, H, H, Cin), device='cuda'); dz = torch.randn((B, H, H, Cout), device='cuda')
    w = torch.randn((k, k, Cin, Cout), device='cuda') * 0.05; bias = torch.zeros(Cout, device='cuda')
    y = torch.empty((B, H, H, Cout), device='cuda'); dw = torch.empty_like(w); db = torch.empty(Cout, device='cuda')
    dx = torch.empty_like(x)
    gf = 2.0 * B * H * H * k * k * Cin * Cout / 1e9
    def fprop(): ssdb.check(L.ssdb_op_conv_fprop(2, P(x), P(w), P(bias), B, H, H, Cin, Cout, k, 1, 1, 1, 1, H, H, 1, P(y), None))
    def dgrad(): ssdb.check(L.ssdb_op_conv_dgrad(2, P(dz), P(w), None, B, H, H, Cin, Cout, k, 1, 1, 1, 1, H, H, 0, P(dx), None))
    def wgrad(): ssdb.check(L.ssdb_op_conv_wgrad(2, P(x), P(dz), B, H, H, Cin, Cout, k, 1, 1, 1, 1, H, H, P(dw), P(db), None))
    os.environ.pop('SSDB_WG_DEBUG', None)
    tf = timeit(fprop); td = timeit(dgrad); tw = timeit(wgrad)
    line = 'B%d H%d %d->%d: GF %.0f | fprop %.3f ms (%.0f TF/s) dgrad %.3f  wgrad %.3f ms (%.0f TF/s)' % (B, H, Cin, Cout, gf, tf, gf / tf, td, tw, gf / tw)
    for dbg in ():
        os.environ['SSDB_WG_DEBUG'] = str(dbg)
        line += ' | dbg%d %.3f' % (dbg, timeit(wgrad))
    os.environ.pop('SSDB_WG_DEBUG', None)
    print(line, flush=True)
print('note: hook timings include the tf32 rounding copies of the operands and (fprop) the filter pack')
