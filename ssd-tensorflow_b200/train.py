#!/usr/bin/env python
"""Thin trainer with the reference's flags (train.py:54-80) on the B200 engine.

The reference's data pipeline (training_data.py, transforms.py, multiprocessing workers) is out of scope
(SURVEY.md section 2); this driver feeds seeded synthetic batches (--synthetic, the default) through the same
call sequence as train.py:166-343: build_from_vgg -> build_optimizer -> sess.run([result, losses, optimizer])
per batch -> decode + NMS of the predictions -> APCalculator (train.py:275-281: from the second epoch on)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ssdutils   # noqa: E402
from average_precision import APCalculator, APs2mAP   # noqa: E402
import synth      # noqa: E402
from ssdvgg import SSDVGG, GlobalStep, Session, piecewise_constant   # noqa: E402
from utils import str2bool   # noqa: E402


def main():
    ap = argparse.ArgumentParser(description='Train the SSD')
    ap.add_argument('--name', default='test')
    ap.add_argument('--data-dir', default='pascal-voc')
    ap.add_argument('--vgg-dir', default='vgg_graph')
    ap.add_argument('--epochs', type=int, default=200)
    ap.add_argument('--batch-size', type=int, default=8)
    ap.add_argument('--tensorboard-dir', default='tb')
    ap.add_argument('--checkpoint-interval', type=int, default=5)
    ap.add_argument('--lr-values', default='0.00075;0.0001;0.00001')
    ap.add_argument('--lr-boundaries', default='320000;400000')
    ap.add_argument('--momentum', type=float, default=0.9)
    ap.add_argument('--weight-decay', type=float, default=0.0005)
    ap.add_argument('--continue-training', type=str2bool, default='False')
    ap.add_argument('--num-workers', type=int, default=os.cpu_count())
    ap.add_argument('--preset', default='vgg300')
    ap.add_argument('--synthetic', type=str2bool, default='True')
    ap.add_argument('--batches-per-epoch', type=int, default=4)
    ap.add_argument('--valid-batches', type=int, default=1, help='validation batches per epoch (train.py:287-303)')
    ap.add_argument('--feed', default='gt', choices=['gt', 'labels'],
                    help="'gt': raw ground-truth boxes, anchors matched inside the fused loss kernels; "
                         "'labels': the reference's dense label tensor (create_labels = LabelCreatorTransform on the GPU)")
    args = ap.parse_args()
    if not args.synthetic:
        print('[!] only --synthetic input is built here: the VOC loader / augmentation pipeline (training_data.py, transforms.py) '
              'is outside the hot path; feed your own batches through Session.run like the loop below does')
        return 1
    preset = ssdutils.get_preset_by_name(args.preset)
    anchors = ssdutils.get_anchors_for_preset(preset)
    lr_values = [float(v) for v in args.lr_values.split(';')]
    lr_boundaries = [int(v) for v in args.lr_boundaries.split(';')]
    with Session() as sess:
        net = SSDVGG(sess, preset)
        ckpt = os.path.join(args.name, 'final.npz')
        step = GlobalStep(0)
        lr = piecewise_constant(step, lr_boundaries, lr_values)
        start_epoch = 0
        if args.continue_training:
            # train.py:101-134: resume from the newest e<N> checkpoint (weights, Momentum slots, global_step, epoch)
            last = _latest_checkpoint(args.name)
            if last is None:
                print('[!] No checkpoints found in ' + args.name)
                return 1
            net.build_from_metagraph(None, last)
            net.build_optimizer_from_metagraph(lr, args.weight_decay, args.momentum, step)
            start_epoch = net.epoch
            print('[i] resuming from %s: epoch %d, global step %d' % (last, start_epoch, step.value))
        else:
            net.build_from_vgg(args.vgg_dir, 20)
            net.build_optimizer(learning_rate=lr, weight_decay=args.weight_decay, momentum=args.momentum, global_step=step)
        side = preset.image_size.w
        lid2name = {i: 'class%d' % i for i in range(20)}
        train_ap, valid_ap = APCalculator(), APCalculator()

        def batch(first):
            x = synth.images(first, args.batch_size, side)
            gts = [synth.gt_boxes(first + i) for i in range(args.batch_size)]
            if args.feed == 'gt':
                gt, cnt = synth.pack_gt(gts, 8)
                return x, gts, {net.image_input: x, net.gt_boxes: gt, net.gt_counts: cnt}
            boxes = [[ssdutils.Box(None, int(g[0]), ssdutils.Point(g[1], g[2]), ssdutils.Size(g[3], g[4])) for g in gt] for gt in gts]
            y, _ = ssdutils.create_labels(boxes, anchors, 20)
            return x, gts, {net.image_input: x, net.labels: y}

        valid_first = 10 ** 6           # a disjoint range of synthetic sample indices plays the validation split
        losses = vlosses = None
        for e in range(start_epoch, args.epochs):
            t0 = time.time()
            train_ap.clear(); valid_ap.clear()
            for b in range(args.batches_per_epoch):          # train.py:254-281
                x, gts, feed = batch((e * args.batches_per_epoch + b) * args.batch_size)
                result, losses, _ = sess.run([net.result, net.losses, net.optimizer], feed_dict=feed)
                if np.isnan(losses['confidence']):
                    print('[!] Confidence loss is NaN.')
                if e == 0:
                    continue
                # train.py:275-278: decode + NMS of every sample, fed to the AP calculator -- one batched launch here
                dets, counts = ssdutils.detect_batch_rows(result, anchors, 0.5, 200)
                train_ap.add_detections_batch(gts, dets, counts, lid2name)
            for b in range(args.valid_batches):              # train.py:287-303: forward + loss only, then decode + NMS + AP
                x, gts, feed = batch(valid_first + b * args.batch_size)
                result, vlosses = sess.run([net.result, net.losses], feed_dict=feed)
                dets, counts = ssdutils.detect_batch_rows(result, anchors, 0.5, 200)
                valid_ap.add_detections_batch(gts, dets, counts, lid2name)
            aps = train_ap.compute_aps() if e > 0 else {}
            vaps = valid_ap.compute_aps() if args.valid_batches else {}
            print('[i] epoch %d: train total %.4f loc %.4f conf %.4f l2 %.4f mAP %.4f | valid total %.4f mAP %.4f | %.2fs' %
                  (e, losses['total'], losses['localization'], losses['confidence'], losses['l2'], APs2mAP(aps),
                   vlosses['total'] if vlosses else float('nan'), APs2mAP(vaps), time.time() - t0))
            net.epoch = e + 1
            if (e + 1) % args.checkpoint_interval == 0:
                os.makedirs(args.name, exist_ok=True)
                net.save(os.path.join(args.name, 'e%d' % (e + 1)))
        os.makedirs(args.name, exist_ok=True)
        net.save(ckpt)
    return 0


def _latest_checkpoint(directory):
    """Highest-numbered e<N>.npz of a run directory (train.py:104-118), or final.npz, or None."""
    import re
    best, path = -1, None
    if os.path.isdir(directory):
        for f in os.listdir(directory):
            m = re.match(r'^e(\d+)\.npz$', f)
            if m and int(m.group(1)) > best:
                best, path = int(m.group(1)), os.path.join(directory, f)
        if path is None and os.path.exists(os.path.join(directory, 'final.npz')):
            path = os.path.join(directory, 'final.npz')
    return path


if __name__ == '__main__':
    sys.exit(main())
