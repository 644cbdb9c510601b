"""Small fixed workload for ncu captures: two training steps of vgg300 at batch 16 (no host work in between)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ssd-tensorflow_b200'))
import ssdb     # noqa: E402
import ssdvgg   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
net = ssdb.Net('vgg300', 20, max_batch=B)
m = ssdvgg.SSDVGG(ssdvgg.Session(), 'vgg300')
P = m._initial_params(20, seed=7)
for k, shape in net.tensors():
    net.set_tensor(k, P[k])
A = net.num_anchors
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.rand((B, 300, 300, 3), device='cuda', generator=g) * 255
labels = torch.zeros((B, A, 25), device='cuda'); labels[..., 20] = 1
idx = torch.randint(0, A, (B, 40), device='cuda', generator=g)
for b in range(B):
    labels[b, idx[b], 20] = 0; labels[b, idx[b], 3] = 1
losses = torch.zeros(4, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for _ in range(steps):
    net.train_step(x.data_ptr(), B, labels_ptr=labels.data_ptr(), lr=1e-9, losses_ptr=losses.data_ptr(), stream=st)
torch.cuda.synchronize()
print('losses', losses.cpu().tolist(), 'launches', ssdb.launch_count())
