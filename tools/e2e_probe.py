"""Where does the host-fed step spend its time?  Times the variants of the host entry points (wall clock, 5 steps each)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ssd-tensorflow_b200')); sys.path.insert(0, ROOT)
import ssdb, ssdvgg, synth
from bench import labels_for, synth_gt
from ssdutils import anchors_as_array, get_anchors_for_preset, get_preset_by_name
preset = sys.argv[1] if len(sys.argv) > 1 else 'vgg300'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
p = get_preset_by_name(preset); side = p.image_size.w
anchors = anchors_as_array(get_anchors_for_preset(p))
sess = ssdvgg.Session(); m = ssdvgg.SSDVGG(sess, p); m.build_from_vgg(None, 20); m.build_optimizer(1e-9, 0.0005, 0.9)
eng = m._ensure_engine(B)
x = torch.from_numpy(synth.images(0, B, side)).pin_memory().numpy()
y = torch.from_numpy(labels_for(0, B, anchors)).pin_memory().numpy()
gt, cnt = synth_gt(0, B)
fixed = torch.empty((B, eng.num_anchors, 25)).pin_memory().numpy()
def t(name, fn, n=5):
    fn(); fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n * 1e3
    print('%-44s %8.2f ms   pool blocks %s' % (name, dt, len(eng._pinned_result.blocks) if eng._pinned_result else 0), flush=True)
xd = torch.from_numpy(x).cuda(); yd = torch.from_numpy(y).cuda(); ld = torch.zeros(4, device='cuda')
st = torch.cuda.current_stream().cuda_stream
t('device step (train_step)', lambda: eng.train_step(xd.data_ptr(), B, labels_ptr=yd.data_ptr(), lr=1e-9, losses_ptr=ld.data_ptr(), stream=st))
t('host labels, fixed pinned result', lambda: eng.train_step_host(x, y, 1e-9, 0.9, 0.0005, result_out=fixed))
t('host labels, no result', lambda: eng.train_step_host(x, y, 1e-9, 0.9, 0.0005, want_result=False))
t('host labels, pooled result', lambda: eng.train_step_host(x, y, 1e-9, 0.9, 0.0005))
t('host gt, fixed pinned result', lambda: eng.train_step_host_gt(x, gt, cnt, 1e-9, 0.9, 0.0005, result_out=fixed))
t('host gt, pooled result', lambda: eng.train_step_host_gt(x, gt, cnt, 1e-9, 0.9, 0.0005))
t('Session.run labels', lambda: sess.run([m.result, m.losses, m.optimizer], feed_dict={m.image_input: x, m.labels: y}))
t('Session.run gt', lambda: sess.run([m.result, m.losses, m.optimizer], feed_dict={m.image_input: x, m.gt_boxes: gt, m.gt_counts: cnt}))
t('forward_host pooled', lambda: eng.forward_host(x))
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
yt = torch.from_numpy(y); e0.record(); yd.copy_(yt, non_blocking=True); e1.record(); torch.cuda.synchronize()
print('H2D labels %.1f MB in %.2f ms' % (y.nbytes / 1e6, e0.elapsed_time(e1)))
