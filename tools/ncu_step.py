"""One profiled training step for ncu (--profile-from-start off): warm-up steps, then cudaProfilerStart .. Stop around one
step.  python tools/ncu_step.py [preset] [B] [mode]   (mode: train | fwd)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ssd-tensorflow_b200'))
import ssdb, ssdvgg
preset = sys.argv[1] if len(sys.argv) > 1 else 'vgg300'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
mode = sys.argv[3] if len(sys.argv) > 3 else 'train'
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
def step():
    if mode == 'fwd':
        net.forward(x.data_ptr(), B, None, st)
    else:
        net.train_step(x.data_ptr(), B, labels_ptr=labels.data_ptr(), lr=1e-9, losses_ptr=losses.data_ptr(), stream=st)
for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('launches', ssdb.launch_count())
