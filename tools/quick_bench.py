"""Quick on-GPU timing of the engine (bring-up aid; bench.py is the judged harness)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ssd-tensorflow_b200'))
import ssdb     # noqa: E402
import ssdvgg   # noqa: E402
from ssdutils import get_preset_by_name  # noqa: E402

preset = sys.argv[1] if len(sys.argv) > 1 else 'vgg300'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
side = 300 if preset == 'vgg300' else 512
net = ssdb.Net(preset, 20, max_batch=B)
m = ssdvgg.SSDVGG(ssdvgg.Session(), preset)
P = m._initial_params(20, seed=7)
for k, shape in net.tensors():
    net.set_tensor(k, P[k])
A = net.num_anchors
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.rand((B, side, side, 3), device='cuda', generator=g) * 255
labels = torch.zeros((B, A, 25), device='cuda'); labels[..., 20] = 1
idx = torch.randint(0, A, (B, 40), device='cuda', generator=g)
for b in range(B):
    labels[b, idx[b], 20] = 0; labels[b, idx[b], 3] = 1
losses = torch.zeros(4, device='cuda')
st = torch.cuda.current_stream().cuda_stream
out = {}
def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
fwd = timeit(lambda: net.forward(x.data_ptr(), B, None, st))
trn = timeit(lambda: net.train_step(x.data_ptr(), B, labels_ptr=labels.data_ptr(), lr=1e-9, losses_ptr=losses.data_ptr(), stream=st))
out['preset'] = preset; out['B'] = B
out['fwd_ms'] = fwd; out['fwd_img_s'] = B / fwd * 1e3
out['train_ms'] = trn; out['train_img_s'] = B / trn * 1e3
out['losses'] = losses.cpu().tolist()
prof = net.profile_step(x.data_ptr(), labels.data_ptr(), B)
out['profile'] = prof
agg = {}
for label, ms, l in prof:
    k = label.split(':')[0]
    agg[k] = agg.get(k, 0.0) + ms
out['profile_by_phase'] = agg
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'quick_bench_%s_%d.json' % (preset, B)), 'w'), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != 'profile'}))
top = sorted(prof, key=lambda t: -t[1])[:25]
for t in top: print('%-40s %8.3f ms  launches %d' % t)
