"""Bring-up experiment: sensitivity of the wgrad kernel to the TMA pixel-box shape (SSDB_WG_BOX)."""
import os, sys, ctypes
import torch
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
cases = [((64, 38, 512, 512, 3), ['2,4,4', '2,2,8', '4,1,8', '1,1,32', '2,8,2', '8,4,1']),
         ((64, 300, 64, 64, 3), ['4,4,4', '4,2,8', '2,2,16', '8,8,1', '4,4,2', '4,2,2', '10,4,1', '20,2,1']),
         ((64, 75, 256, 256, 3), ['8,4,1', '1,4,8', '5,1,8', '1,1,32', '4,4,2'])]
for (B, H, Cin, Cout, k), boxes in cases:
    x = torch.randn((B, H, H, Cin), device='cuda'); dz = torch.randn((B, H, H, Cout), device='cuda')
    dw = torch.empty((k, k, Cin, Cout), device='cuda'); db = torch.empty(Cout, device='cuda')
    gf = 2.0 * B * H * H * k * k * Cin * Cout / 1e9
    def wgrad(): ssdb.check(L.ssdb_op_conv_wgrad(2, P(x), P(dz), B, H, H, Cin, Cout, k, 1, 1, 1, 1, H, H, P(dw), P(db), None))
    os.environ.pop('SSDB_WG_BOX', None)
    line = 'B%d H%d %d->%d GF %.0f: auto %.3f ms' % (B, H, Cin, Cout, gf, timeit(wgrad))
    for bx in boxes:
        os.environ['SSDB_WG_BOX'] = bx
        line += ' | %s %.3f' % (bx, timeit(wgrad))
    os.environ.pop('SSDB_WG_BOX', None)
    print(line, flush=True)
