"""N-rank data-parallel step == 1-rank step on the concatenated batch, on real GPUs.
Every rank runs the N-rank step on its shard; rank 0 then runs the full batch alone and compares parameters."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ssd-tensorflow_b200'))
import ssdb, ssdvgg, synth
from parallel import DataParallelTrainer, shard_range
from ssdutils import anchors_as_array, get_anchors_for_preset, get_preset_by_name

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl')
preset = get_preset_by_name('vgg300')
anchors = anchors_as_array(get_anchors_for_preset(preset))
G = 4 * world                      # global batch
lo, hi = shard_range(G, rank, world)
m = ssdvgg.SSDVGG(ssdvgg.Session(), preset)
P = m._initial_params(20, seed=7)
def labels_for(first, count):
    gts = [synth.gt_boxes(first + i) for i in range(count)]
    gt, cnt = synth.pack_gt(gts, 8)
    return ssdb.match_anchors_host(gt, cnt, anchors, 20, want_match=False)[1]
def run(batch_lo, batch_hi, trainer_world, host=False):
    net = ssdb.Net('vgg300', 20, max_batch=batch_hi - batch_lo)
    for k, shape in net.tensors():
        net.set_tensor(k, P[k])
    x = torch.from_numpy(synth.images(batch_lo, batch_hi - batch_lo, 300)).cuda()
    y = torch.from_numpy(labels_for(batch_lo, batch_hi - batch_lo)).cuda()
    if trainer_world > 1 and host:
        # the host-fed path with raw ground truth: begin -> bucketed all-reduce + update on the side stream -> end
        tr = DataParallelTrainer(net)
        gt, cnt = synth.pack_gt([synth.gt_boxes(batch_lo + i) for i in range(batch_hi - batch_lo)], 8)
        tr.step_host_gt(x.cpu().numpy(), gt, cnt, 0.00075, 0.9, 0.0005)
    elif trainer_world > 1:
        tr = DataParallelTrainer(net)
        tr.step(x.data_ptr(), y.data_ptr(), batch_hi - batch_lo, 0.00075, 0.9, 0.0005)
    else:
        st = torch.cuda.current_stream().cuda_stream
        net.train_step(x.data_ptr(), batch_hi - batch_lo, labels_ptr=y.data_ptr(), lr=0.00075, momentum=0.9, weight_decay=0.0005, stream=st)
    torch.cuda.synchronize()
    out = {k: net.get_tensor(k, shape) for k, shape in net.tensors()}
    net.close()
    return out
dp = run(lo, hi, world)
dist.barrier()
dp_host = run(lo, hi, world, host=True)
dist.barrier()
if rank == 0:
    single = run(0, G, 1)
    ok = True
    for tag, got in (('device-fed step (bucketed all-reduce behind the backward)', dp), ('host-fed gt step (begin / end)', dp_host)):
        worst = 0.0; name = None
        for k in got:
            step = np.abs(single[k] - P[k]).max()
            e = np.abs(got[k] - single[k]).max() / max(step, 1e-12)
            if e > worst: worst, name = e, k
        print('DP parity, %s: world %d, global batch %d: worst |w_dp - w_single| / |update| = %.3e (%s)' % (tag, world, G, worst, name))
        ok = ok and worst < 0.05
    same = max(float(np.abs(dp[k] - dp_host[k]).max()) for k in dp)
    print('device-fed vs host-fed DP step: max |difference| = %.3e' % same)
    print('PASS' if ok else 'FAIL')
dist.barrier()
dist.destroy_process_group()
