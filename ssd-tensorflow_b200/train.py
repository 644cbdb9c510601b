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
    args = ap.parse_args()
    if not args.synthetic:
        print('[!] only --synthetic input is built here (the VOC loader is outside the hot path)')
        return 1
    preset = ssdutils.get_preset_by_name(args.preset)
    anchors = ssdutils.get_anchors_for_preset(preset)
    lr_values = [float(v) for v in args.lr_values.split(';')]
    lr_boundaries = [int(v) for v in args.lr_boundaries.split(';')]
    with Session() as sess:
        net = SSDVGG(sess, preset)
        ckpt = os.path.join(args.name, 'final.npz')
        if args.continue_training and os.path.exists(ckpt):
            net.build_from_metagraph(None, ckpt)
        else:
            net.build_from_vgg(args.vgg_dir, 20)
        step = GlobalStep(0)
        net.build_optimizer(learning_rate=piecewise_constant(step, lr_boundaries, lr_values),
                            weight_decay=args.weight_decay, momentum=args.momentum, global_step=step)
        side = preset.image_size.w
        lid2name = {i: 'class%d' % i for i in range(20)}
        ap_calc = APCalculator()
        for e in range(args.epochs):
            t0 = time.time()
            ap_calc.clear()
            for b in range(args.batches_per_epoch):
                first = (e * args.batches_per_epoch + b) * args.batch_size
                x = synth.images(first, args.batch_size, side)
                gts = [synth.gt_boxes(first + i) for i in range(args.batch_size)]
                boxes = [[ssdutils.Box(None, int(g[0]), ssdutils.Point(g[1], g[2]), ssdutils.Size(g[3], g[4])) for g in gt] for gt in gts]
                y, _ = ssdutils.create_labels(boxes, anchors, 20)
                result, losses, _ = sess.run([net.result, net.losses, net.optimizer],
                                             feed_dict={net.image_input: x, net.labels: y})
                if np.isnan(losses['confidence']):
                    print('[!] Confidence loss is NaN.')
                if e == 0:
                    continue
                # train.py:275-278: decode + NMS of every sample, fed to the AP calculator -- one batched launch here
                dets, counts = ssdutils.detect_batch_rows(result, anchors, 0.5, 200)
                ap_calc.add_detections_batch(gts, dets, counts, lid2name)
            aps = ap_calc.compute_aps() if e > 0 else {}
            print('[i] epoch %d: total %.4f loc %.4f conf %.4f l2 %.4f | mAP %.4f over %d classes | %.2fs' %
                  (e, losses['total'], losses['localization'], losses['confidence'], losses['l2'],
                   APs2mAP(aps), len(aps), time.time() - t0))
            if (e + 1) % args.checkpoint_interval == 0:
                os.makedirs(args.name, exist_ok=True)
                net.save(os.path.join(args.name, 'e%d' % (e + 1)))
        os.makedirs(args.name, exist_ok=True)
        net.save(ckpt)
    return 0


if __name__ == '__main__':
    sys.exit(main())
