"""Fixed workload for ncu captures of the HBM-bound kernels: dense + fused-match multibox loss at batch 64
and decode + NMS at batch 128 (BASELINE.json configs[1] / configs[4] sizes), stateless C-ABI entry points."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ssd-tensorflow_b200'))
import ssdb      # noqa: E402
import synth     # noqa: E402
from ssdutils import anchors_as_array, get_anchors_for_preset, get_preset_by_name   # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
anc = anchors_as_array(get_anchors_for_preset(get_preset_by_name('vgg300')))
A = anc.shape[0]
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ad = torch.from_numpy(anc).cuda()

B = 64
gts = [synth.gt_boxes(i) for i in range(B)]
gt, cnt = synth.pack_gt(gts, 8)
_, labels = ssdb.match_anchors_host(gt, cnt, anc, 20, want_match=False)
out = torch.randn((B, A, 25), device='cuda') * 2
ld = torch.from_numpy(labels).cuda(); gd = torch.from_numpy(gt).cuda(); cd = torch.from_numpy(cnt).cuda()
g = torch.empty_like(out); r = torch.empty_like(out); l = torch.zeros(2, device='cuda')
for _ in range(reps):
    ssdb.check(ssdb.lib().ssdb_multibox_loss(P(out), P(ld), B, A, 20, 1.0, P(l), P(g), P(r), st))
    ssdb.check(ssdb.lib().ssdb_multibox_loss_gt(P(out), P(gd), P(cd), B, 8, P(ad), A, 20, 1.0, P(l), P(g), P(r), None, st))
torch.cuda.synchronize()
print('loss', l.cpu().tolist())

NB = 128
pred = torch.from_numpy(np.stack([synth.pred_clustered(1000 + i, anc) for i in range(NB)])).cuda()
dets = torch.zeros((NB, 200, 8), dtype=torch.int32, device='cuda'); counts = torch.zeros((NB, 2), dtype=torch.int32, device='cuda')
for _ in range(reps):
    ssdb.check(ssdb.lib().ssdb_decode_nms(P(pred), NB, A, 20, P(ad), 0.01, 200, 0.45, P(dets), P(counts), st))
torch.cuda.synchronize()
print('nms kept', int(counts[:, 0].sum()), 'candidates', int(counts[:, 1].sum()), 'launches', ssdb.launch_count())
