"""Per-call GPU (CUDA events) and host (perf_counter) time of ssdb_decode_nms, v1 vs v2, rotating vs single buffer."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ssd-tensorflow_b200'))
import ssdb      # noqa: E402
import synth     # noqa: E402
from ssdutils import anchors_as_array, get_anchors_for_preset, get_preset_by_name   # noqa: E402

anc = anchors_as_array(get_anchors_for_preset(get_preset_by_name('vgg300')))
A = anc.shape[0]
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ad = torch.from_numpy(anc).cuda()
NB = 128
pred = np.stack([synth.pred_clustered(1000 + i, anc) for i in range(NB)])
pds = [torch.from_numpy(pred).cuda() for _ in range(3)]
dets = torch.zeros((NB, 200, 8), dtype=torch.int32, device='cuda'); counts = torch.zeros((NB, 2), dtype=torch.int32, device='cuda')
for impl in ('v2', 'v1', 'v2'):
    os.environ['SSDB_NMS'] = impl
    for rot in (3, 1):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(12)]
        host = []
        for i in range(3):
            ssdb.check(ssdb.lib().ssdb_decode_nms(P(pds[i % rot]), NB, A, 20, P(ad), 0.01, 200, 0.45, P(dets), P(counts), st))
        torch.cuda.synchronize()
        t_all = time.perf_counter()
        for i, (a, b) in enumerate(evs):
            t0 = time.perf_counter()
            a.record()
            ssdb.check(ssdb.lib().ssdb_decode_nms(P(pds[i % rot]), NB, A, 20, P(ad), 0.01, 200, 0.45, P(dets), P(counts), st))
            b.record()
            host.append((time.perf_counter() - t0) * 1e6)
        torch.cuda.synchronize()
        t_all = (time.perf_counter() - t_all) * 1e6 / len(evs)
        gpu = [a.elapsed_time(b) * 1e3 for a, b in evs]
        print(impl, 'buffers', rot, 'gpu us', [round(g) for g in gpu], 'host us', [round(h) for h in host], 'wall/call us', round(t_all),
              'kept', int(counts[:, 0].sum()))
