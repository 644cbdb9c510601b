#!/usr/bin/env python
"""Thin inference driver (reference infer.py:61-264): restore a model, run batches through
sess.run(net.result, {image_input: x, keep_prob: 1}), decode + suppress on the GPU, then the reference's
post-NMS steps: --compute-stats (APCalculator, infer.py:260,273-279), --pascal-summary (infer.py:262-264),
--dump-predictions (infer.py:251-258).  Input is synthetic (the VOC loader is outside the hot path)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ssdutils   # noqa: E402
import synth      # noqa: E402
from average_precision import APCalculator, APs2mAP   # noqa: E402
from pascal_summary import PascalSummary   # noqa: E402
from ssdvgg import SSDVGG, Session   # noqa: E402
from utils import Size, str2bool   # noqa: E402


def main():
    ap = argparse.ArgumentParser(description='SSD inference')
    ap.add_argument('--name', default='test')
    ap.add_argument('--checkpoint', default='')
    ap.add_argument('--preset', default='vgg300')
    ap.add_argument('--threshold', type=float, default=0.5)
    ap.add_argument('--batch-size', type=int, default=32)
    ap.add_argument('--batches', type=int, default=1)
    ap.add_argument('--output-dir', default='test-output')
    ap.add_argument('--dump-predictions', type=str2bool, default='False')
    ap.add_argument('--compute-stats', type=str2bool, default='True')
    ap.add_argument('--pascal-summary', type=str2bool, default='False')
    args = ap.parse_args()
    preset = ssdutils.get_preset_by_name(args.preset)
    anchors = ssdutils.get_anchors_for_preset(preset)
    with Session() as sess:
        net = SSDVGG(sess, preset)
        ckpt = args.checkpoint or os.path.join(args.name, 'final.npz')
        if os.path.exists(ckpt):
            net.build_from_metagraph(None, ckpt)
        else:
            print('[!] no checkpoint at %s: using freshly initialised weights' % ckpt)
            net.build_from_vgg(None, 20)
        lid2name = {i: 'class%d' % i for i in range(20)}
        ap_calc = APCalculator() if args.compute_stats else None
        pascal = PascalSummary() if args.pascal_summary else None
        if args.dump_predictions or args.pascal_summary:
            os.makedirs(args.output_dir, exist_ok=True)
        side = preset.image_size.w
        for b in range(args.batches):
            first = b * args.batch_size
            x = synth.images(first, args.batch_size, side)
            # infer.py:225-235 as one flow on the device: forward, then decode (no cap) + suppress on the resident result
            # tensor; only the detections cross PCIe (the 873 KB/image result comes back only for --dump-predictions).
            # Then, like the reference, keep the first 200 of the class-grouped list.
            if args.dump_predictions:
                result = sess.run(net.result, feed_dict={net.image_input: x, net.keep_prob: 1})
                dets = ssdutils.detect_batch(result, anchors, args.threshold, lid2name, None)
            else:
                dets = net.detect(x, args.threshold, lid2name, None)
            dets = [d[:200] for d in dets]
            print('[i] batch %d: %s detections per image' % (b, [len(d) for d in dets][:8]))
            for i, boxes in enumerate(dets):
                gt = synth.gt_boxes(first + i)
                gt_boxes = [ssdutils.Box(lid2name[int(g[0])], int(g[0]), ssdutils.Point(g[1], g[2]), ssdutils.Size(g[3], g[4])) for g in gt]
                if ap_calc:
                    ap_calc.add_detections(gt_boxes, boxes)
                if pascal:
                    pascal.add_detections('synthetic_%06d.jpg' % (first + i), boxes, Size(side, side))
                if args.dump_predictions:
                    np.save(os.path.join(args.output_dir, 'synthetic_%06d.npy' % (first + i)), result[i])
        if ap_calc:
            aps = ap_calc.compute_aps()
            for k in sorted(aps):
                print('[i] AP [%s]: %.3f' % (k, aps[k]))
            print('[i] mAP: %.3f' % APs2mAP(aps))
        if pascal:
            pascal.write_summary(args.output_dir)
    return 0


if __name__ == '__main__':
    sys.exit(main())
