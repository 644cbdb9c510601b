"""Is decode_nms_kernel bound by cold instruction fetch?  Time batches of 128 .. 1184 images (1 .. 8 CTAs per SM) and trace a late CTA."""
import ctypes, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ssd-tensorflow_b200'))
import ssdb, synth
from ssdutils import anchors_as_array, get_anchors_for_preset, get_preset_by_name
anc = anchors_as_array(get_anchors_for_preset(get_preset_by_name('vgg300')))
A = anc.shape[0]
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ad = torch.from_numpy(anc).cuda()
base = torch.from_numpy(np.stack([synth.pred_clustered(1000 + i, anc) for i in range(128)])).cuda()
for NB in (128, 148, 296, 592, 1184):
    pred = base.repeat((NB + 127) // 128, 1, 1)[:NB].contiguous()
    dets = torch.zeros((NB, 200, 8), dtype=torch.int32, device='cuda'); counts = torch.zeros((NB, 2), dtype=torch.int32, device='cuda')
    for impl in ('v2', 'v1'):
        os.environ['SSDB_NMS'] = impl
        for _ in range(2):
            ssdb.check(ssdb.lib().ssdb_decode_nms(P(pred), NB, A, 20, P(ad), 0.01, 200, 0.45, P(dets), P(counts), st))
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ssdb.check(ssdb.lib().ssdb_decode_nms(P(pred), NB, A, 20, P(ad), 0.01, 200, 0.45, P(dets), P(counts), st))
        e1.record(); torch.cuda.synchronize()
        print('images', NB, impl, 'us/call', round(e0.elapsed_time(e1) * 100, 1), 'kept', int(counts[:, 0].sum()), flush=True)
os.environ['SSDB_NMS'] = 'v2'
os.environ['SSDB_TRACE'] = '1'
for blk in (0, 700, 1100):
    os.environ['SSDB_TRACE_BLOCK'] = str(blk)
    print('trace of CTA', blk, flush=True)
    ssdb.check(ssdb.lib().ssdb_decode_nms(P(pred), NB, A, 20, P(ad), 0.01, 200, 0.45, P(dets), P(counts), st))
    torch.cuda.synchronize()
