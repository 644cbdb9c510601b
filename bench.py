#!/usr/bin/env python
"""Benchmark of the SSD-VGG hot path on B200 (contract in the task brief).

    python bench.py --gpus N --steps K --warmup W            # this repo's engine
    python bench.py --impl reference --gpus N ...            # the reference's path on the host cores

A "step" is one pass of the hot path over one synthetic batch: vgg300, 64 images per GPU,
forward + multibox loss + backward + Momentum update (BASELINE.json configs[1]; at N = 8 the
global batch is 512 = configs[2]).  `value` is whole-job images/s with inputs resident in HBM;
`e2e` is the same step through the reference-facing call (SSDVGG / Session.run at N = 1, the
data-parallel trainer at N > 1) with HOST buffers: H2D of images + labels and D2H of the result
and losses inside the timed region.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, 'ssd-tensorflow_b200'))

METRIC = 'images/sec vgg300 fwd+bwd+loss'
FWD_GFLOP = {'vgg300': 62.747, 'vgg512': 180.415}            # SURVEY.md App. B, per image
TRAIN_GFLOP = {'vgg300': 187.93, 'vgg512': 540.34}
FALLBACK_PEAKS = {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}  # B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--preset', default='vgg300', choices=['vgg300', 'vgg512'])
    ap.add_argument('--batch', type=int, default=0, help='images per GPU (default 64 for vgg300, 32 for vgg512)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return d, 'measured'
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), 'fallback'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=3)
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith('active') for r in self.rows if len(r) > 3 + i)]
        try:
            mx = float(self.rows[0][1])
        except Exception:
            mx = None
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': reasons, 'samples': len(self.rows)}


def ncu_traffic():
    """Average DRAM bytes per conv launch from the committed `ncu --set full` summary (profiles/), or None."""
    path = os.path.join(ROOT, 'profiles', 'r1_ncu_conv_tc_b64.txt')
    if not os.path.exists(path):
        return None, None
    mult = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    tot, n = 0.0, 0
    for line in open(path):
        parts = line.split()
        if len(parts) >= 3 and parts[0] in ('dram__bytes_read.sum', 'dram__bytes_write.sum') and parts[2] in mult:
            tot += float(parts[1]) * mult[parts[2]]
            n += parts[0] == 'dram__bytes_read.sum'
    return (tot / n, n) if n else (None, None)


def ncu_box_traffic(kernel_prefixes):
    """Sum of dram__bytes_read + dram__bytes_write of the FIRST launch of each named kernel in the committed `ncu --set full`
    summary of the loss / decode+NMS kernels (profiles/r1_ncu_box_kernels.txt: 64 images for the loss, 128 for NMS), or None."""
    path = os.path.join(ROOT, 'profiles', 'r1_ncu_box_kernels.txt')
    if not os.path.exists(path):
        return None
    mult = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    seen, cur, tot = set(), None, 0.0
    for line in open(path):
        if line.startswith('---'):
            name = line.split('::')[-1].strip()
            cur = next((k for k in kernel_prefixes if name.startswith(k) and k not in seen), None)
            if cur:
                seen.add(cur)
            continue
        parts = line.split()
        if cur and len(parts) >= 3 and parts[0] in ('dram__bytes_read.sum', 'dram__bytes_write.sum') and parts[2] in mult:
            tot += float(parts[1]) * mult[parts[2]]
    return tot if len(seen) == len(kernel_prefixes) else None


def labels_for(first, count, preset_name, anchors):
    """Dense labels for the synthetic GT boxes, built by the GPU matcher (product path)."""
    import ssdb
    import synth
    gts = [synth.gt_boxes(first + i) for i in range(count)]
    gt, cnt = synth.pack_gt(gts, 8)
    _, labels = ssdb.match_anchors_host(gt, cnt, anchors, 20, want_match=False)
    return labels


def oracle_rate(preset, sample_batch, steps, threads=None):
    """images/s of the torch-CPU restatement of the reference graph (forward+loss+backward+update)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import torch
    import box_oracle as bo
    import net_oracle as no
    import synth
    if threads:
        torch.set_num_threads(threads)
    side = bo.PRESETS[preset]['image']
    anc = bo.anchors(preset)
    aabs = bo.anchors_abs(anc)
    P = no.init_params(preset, dtype=torch.float32)
    V = {k: torch.zeros_like(v) for k, v in P.items()}
    x = torch.tensor(synth.images(0, sample_batch, side))
    y = torch.tensor(np.stack([bo.make_labels(synth.gt_boxes(i), anc, aabs, 20)[0] for i in range(sample_batch)]))
    times = []
    for s in range(steps):
        t0 = time.perf_counter()
        no.train_step(P, V, x, y, preset)
        times.append(time.perf_counter() - t0)
    return sample_batch / float(np.median(times)), float(np.median(times)), torch.get_num_threads()


def run_reference(args):
    """The reference's own implementation of the path, timed on the host cores.  TensorFlow 1.x (the
    reference's runtime) is not installable here, so this is the oracle port (kind 'port')."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    preset = args.preset
    sample = 2
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    rate, sec, cores = oracle_rate(preset, sample, steps + warm)
    line = {
        'impl': 'reference', 'metric': METRIC.replace('vgg300', preset), 'value': rate, 'unit': 'images/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warm, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': '%s forward+loss+backward+update, torch-CPU restatement of ssdvgg.py (TensorFlow 1.x not installable)' % preset,
                   'sample_batch': sample},
        'cpu_baseline': {'value': rate, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d steps of batch %d on the host cores, median' % (steps, sample)},
        'e2e': {'value': rate, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)
    import torch
    import torch.distributed as dist
    import ssdb
    import ssdvgg
    import synth
    from parallel import DataParallelTrainer
    from ssdutils import anchors_as_array, get_anchors_for_preset, get_preset_by_name

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local)
        dist.init_process_group('nccl')
    else:
        torch.cuda.set_device(0)
    ssdb.require_device()
    preset = args.preset
    B = args.batch or (64 if preset == 'vgg300' else 32)
    p = get_preset_by_name(preset)
    side = p.image_size.w
    anchors = anchors_as_array(get_anchors_for_preset(p))
    A = anchors.shape[0]

    # model through the reference-facing surface; the engine handle underneath is shared by both timings
    sess = ssdvgg.Session()
    model = ssdvgg.SSDVGG(sess, p)
    model.build_from_vgg(None, 20)
    step = ssdvgg.GlobalStep(0)
    model.build_optimizer(learning_rate=ssdvgg.piecewise_constant(step, [320000, 400000], [0.00075, 0.0001, 0.00001]),
                          weight_decay=0.0005, momentum=0.9, global_step=step)
    eng = model._ensure_engine(B)
    trainer = DataParallelTrainer(eng)
    trainer.broadcast_parameters(0)

    first = rank * B
    x_host = torch.from_numpy(synth.images(first, B, side)).pin_memory()
    y_host = torch.from_numpy(labels_for(first, B, preset, anchors)).pin_memory()
    x_dev = x_host.cuda()
    y_dev = y_host.cuda()
    losses_dev = torch.zeros(4, device='cuda')
    result_dev = torch.empty((B, A, 25), device='cuda')
    res_host = torch.empty((B, A, 25)).pin_memory()
    st = torch.cuda.current_stream().cuda_stream
    lr, mu, wd = 0.00075, 0.9, 0.0005

    def dev_step():
        trainer.step(x_dev.data_ptr(), y_dev.data_ptr(), B, lr, mu, wd, losses_ptr=losses_dev.data_ptr(), result_ptr=result_dev.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        dev_step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ssdb.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        dev_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ssdb.launch_count() - l0
    clocks = sampler.summary()
    if world > 1:
        t = torch.tensor([ms], device='cuda'); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    ms_step = ms / args.steps
    value = world * B / ms_step * 1e3
    final_losses = losses_dev.cpu().numpy().tolist()

    # ---- end to end: host buffers in, result + losses out, every step
    x_np, y_np = x_host.numpy(), y_host.numpy()
    res_np = res_host.numpy()

    def e2e_step():
        if world == 1:
            # the reference-facing call: sess.run([net.result, net.losses, net.optimizer], feed_dict) (train.py:262-266)
            sess.run([model.result, model.losses, model.optimizer], feed_dict={model.image_input: x_np, model.labels: y_np})
        else:
            trainer.step_host(x_np, y_np, lr, mu, wd)
    if world == 1:
        r, l = sess.run([model.result, model.losses, model.optimizer], feed_dict={model.image_input: x_np, model.labels: y_np})[:2]
        assert r.shape == (B, A, 25) and np.isfinite(l['total'])
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) if world == 1 else e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], device='cuda'); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_ms = float(t.item())
    e2e_value = world * B / (e2e_ms / args.steps) * 1e3
    h2d = x_np.nbytes + y_np.nbytes
    d2h = res_np.nbytes + 16

    # ---- roofline of the dominant kernel family (tcgen05 implicit-GEMM convolutions), rank 0
    roof = None
    if rank == 0:
        prof = eng.profile_step(x_dev.data_ptr(), y_dev.data_ptr(), B)
        pk, pk_kind = peaks()
        tensor_peak = float(pk.get('bf16_tflops_sustained', pk.get('bf16_tflops', FALLBACK_PEAKS['bf16_tflops']))) / 2.0
        conv_ms = sum(ms_ for lab, ms_, _ in prof if lab.split(':')[0] in ('fwd', 'bwd_w', 'bwd_d') and 'pool' not in lab and 'l2_norm' not in lab)
        conv_launch = sum(l_ for lab, _, l_ in prof if lab.split(':')[0] in ('fwd', 'bwd_w', 'bwd_d') and 'pool' not in lab and 'l2_norm' not in lab)
        flops = TRAIN_GFLOP[preset] * 1e9 * B
        achieved = flops / (conv_ms * 1e-3) / 1e12
        by_phase = {}
        for lab, ms_, _ in prof:
            by_phase[lab.split(':')[0]] = by_phase.get(lab.split(':')[0], 0.0) + ms_
        traffic, ncap = ncu_traffic()
        roof = {'bound': 'tensor', 'achieved': achieved, 'peak': tensor_peak, 'unit': 'TFLOP/s', 'frac': achieved / tensor_peak,
                'traffic': traffic, 'traffic_note': 'mean dram__bytes_read+write per launch over the %s conv launches of profiles/r1_ncu_conv_tc_b64.txt '
                                                    '(ncu --set full, batch 64)' % ncap if traffic else None, 'kernel': 'conv_tc_kernel + conv_tc_wgrad_kernel (tcgen05 kind::tf32 implicit GEMM), all conv launches of one step',
                'launches': conv_launch, 'ms_per_step_in_kernel': conv_ms,
                'peak_source': '%s bf16 dense / 2 (tf32 runs at half the bf16 rate)' % pk_kind,
                'step_breakdown_ms': by_phase}

    # ---- the fused multibox loss on its own (HBM-bound kernel family of the step), rank 0
    loss_info = None
    if rank == 0:
        import ctypes
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        out_d = torch.randn((B, A, 25), device='cuda') * 2
        g_d = torch.empty_like(out_d); r_d = torch.empty_like(out_d); l_d = torch.zeros(2, device='cuda')
        def loss_step():
            ssdb.check(ssdb.lib().ssdb_multibox_loss(P(out_d), P(y_dev), B, A, 20, 1.0, P(l_d), P(g_d), P(r_d), ctypes.c_void_p(st)))
        for _ in range(3):
            loss_step()
        torch.cuda.synchronize()
        l0 = ssdb.launch_count()
        e0.record()
        for _ in range(20):
            loss_step()
        e1.record(); torch.cuda.synchronize()
        t_ms = e0.elapsed_time(e1) / 20
        pk, pk_kind = peaks()
        hbm = float(pk.get('hbm_gbs', FALLBACK_PEAKS['hbm_gbs']))
        bytes_alg = 4 * B * A * 25 * 4          # read head output + dense labels, write gradient + net.result
        loss_info = {'kernel': 'loss_rows_kernel + loss_select_kernel + loss_grad_kernel (dense-label multibox loss, batch %d)' % B,
                     'ms': t_ms, 'launches_per_call': (ssdb.launch_count() - l0) // 20,
                     'roofline': {'bound': 'hbm', 'achieved': bytes_alg / (t_ms * 1e-3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
                                  'frac': bytes_alg / (t_ms * 1e-3) / 1e9 / hbm,
                                  'traffic': ncu_box_traffic(['loss_rows_kernel<0', 'loss_select_kernel', 'loss_grad_kernel<0']) if B == 64 else None,
                                  'traffic_note': 'dram bytes of the three kernels of one call, profiles/r1_ncu_box_kernels.txt (ncu --set full, 64 images; '
                                                  'writes still in the L2 when a kernel ends are not counted by ncu)',
                                  'peak_source': pk_kind,
                                  'note': 'algorithmic bytes = 4 x [B,A,25] f32 = %.1f MB (3.49 MB/img); the four tensors (%.0f MB) exceed the L2' % (bytes_alg / 1e6, bytes_alg / 1e6)}}
        del out_d, g_d, r_d

    # ---- second metric of BASELINE.json: batched decode + class-wise NMS (configs[4]), rank 0, device-resident pred
    nms = None
    if rank == 0 and preset == 'vgg300':
        import ctypes
        NB = 128
        pred = np.stack([synth.pred_clustered(1000 + i, anchors) for i in range(NB)])
        # three copies at distinct addresses, used round-robin: 3 x 112 MB > the 126 MB L2, so every timed call reads pred from HBM
        pds = [torch.from_numpy(pred).cuda() for _ in range(3)]
        ad = torch.from_numpy(anchors).cuda()
        dets = torch.zeros((NB, 200, 8), dtype=torch.int32, device='cuda'); cnt = torch.zeros((NB, 2), dtype=torch.int32, device='cuda')
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        def nms_step(i):
            ssdb.check(ssdb.lib().ssdb_decode_nms(P(pds[i % 3]), NB, A, 20, P(ad), 0.01, 200, 0.45, P(dets), P(cnt), ctypes.c_void_p(st)))
        for i in range(3):
            nms_step(i)
        torch.cuda.synchronize()
        l0 = ssdb.launch_count()
        e0.record()
        for i in range(21):
            nms_step(i)
        e1.record(); torch.cuda.synchronize()
        t_ms = e0.elapsed_time(e1) / 21
        nms_launches = (ssdb.launch_count() - l0) // 21
        cands = int(cnt[:, 1].sum().item()); kept = int(cnt[:, 0].sum().item())
        gbs = NB * A * 25 * 4 / (t_ms * 1e-3) / 1e9
        pk, pk_kind = peaks()
        hbm = float(pk.get('hbm_gbs', FALLBACK_PEAKS['hbm_gbs']))
        # end to end: host pred in, host detections out (ssdb_decode_nms_host: H2D of 112 MB + kernels + D2H), as infer.py would call it
        ssdb.decode_nms_host(pred, anchors, 0.01, 200, 0.45)
        t0 = time.perf_counter()
        for _ in range(3):
            ssdb.decode_nms_host(pred, anchors, 0.01, 200, 0.45)
        e2e_nms_ms = (time.perf_counter() - t0) / 3 * 1e3
        # the reference's own NumPy path (restated, pinned against the real code): single thread like train.py:275-278
        nms_cpu = None
        if not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, 'oracle'))
            import box_oracle as bo
            t0 = time.perf_counter(); nimg = 0
            while nimg < 16 and time.perf_counter() - t0 < 10:
                bo.detect(pred[nimg], anchors, 0.01, 200); nimg += 1
            dt = time.perf_counter() - t0
            nms_cpu = {'value': 200 * nimg / dt, 'unit': 'candidate boxes/s', 'images_per_s': nimg / dt, 'cores': 1, 'kind': 'port',
                       'sample': '%d images of the same batch, decode_boxes + suppress_overlaps restated in NumPy (oracle/box_oracle.py), one thread' % nimg}
        nms = {'metric': 'NMS boxes/sec (decode_boxes + class-wise NMS, batch 128, 8732 anchors, cap 200, thr 0.01, IoU 0.45, clustered input)',
               'value': cands / (t_ms * 1e-3), 'unit': 'candidate boxes/s', 'images_per_s': NB / (t_ms * 1e-3), 'ms_per_batch': t_ms,
               'anchors_scanned_per_s': NB * A / (t_ms * 1e-3), 'candidates': cands, 'kept': kept, 'gpu_launches_per_call': nms_launches,
               'l2': 'three pred buffers used round-robin (336 MB > L2): every call streams pred from HBM',
               'e2e': {'value': cands / (e2e_nms_ms * 1e-3), 'unit': 'candidate boxes/s', 'ms_per_batch': e2e_nms_ms,
                       'h2d_bytes_per_step': int(pred.nbytes + anchors.nbytes), 'd2h_bytes_per_step': int(NB * 200 * 8 * 4 + NB * 8),
                       'call': 'ssdb.decode_nms_host(pred, anchors, 0.01, 200, 0.45) -> ssdb_decode_nms_host (pageable host buffers)'},
               'cpu_baseline': nms_cpu,
               'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': gbs / hbm,
                            'traffic': ncu_box_traffic(['decode_scan_kernel', 'decode_nms_kernel']),
                            'traffic_note': 'dram bytes of the two kernels of one call, profiles/r1_ncu_box_kernels.txt (ncu --set full, 128 images)',
                            'note': 'algorithmic bytes = read of pred [128,8732,25] f32 (112 MB) by decode_scan_kernel; the per-image '
                                    'select / sort / greedy-NMS kernel that follows is latency-bound and is inside the timed region',
                            'peak_source': pk_kind}}
        del pds

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            rate, sec, cores = oracle_rate(preset, 2, 2)
            cpu = {'value': rate, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                   'sample': '2 steps of batch 2 (same workload, torch-CPU restatement of the reference graph), median'}
        except Exception as ex:      # the baseline must never take the GPU number down with it
            cpu = {'value': None, 'unit': 'images/s', 'cores': os.cpu_count(), 'kind': 'port', 'sample': 'failed: %r' % (ex,)}

    if rank == 0:
        line = {
            'metric': METRIC.replace('vgg300', preset), 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'tf32', 'data': 'synthetic',
            'config': {'workload': '%s batch %d per GPU, forward+multibox loss+backward+Momentum update (BASELINE.json configs[1]%s)'
                                   % (preset, B, '' if world == 1 else '; global batch %d' % (world * B)),
                       'global_batch': world * B, 'image': side, 'anchors': A, 'parallelism': 'dp%d' % world,
                       'l2': 'no flush: the step streams ~13 GB of activations and gradients, far larger than the 126 MB L2',
                       'operands': 'tf32 tensor-core operands, fp32 accumulate and storage'},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'images/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'ms_per_step': e2e_ms / args.steps,
                    'call': 'Session.run([net.result, net.losses, net.optimizer], feed_dict) -> ssdb_train_step_host' if world == 1
                            else 'DataParallelTrainer.step_host (ssdb_train_step_host_noupdate -> NCCL all-reduce -> ssdb_apply_update)'},
            'gpu_launches': int(launches), 'losses': final_losses,
            'roofline': roof, 'cpu_baseline': cpu, 'loss': loss_info, 'nms': nms,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
