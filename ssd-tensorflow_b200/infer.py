#!/usr/bin/env python
"""Thin inference driver (reference infer.py:61-264): restore a model, run batches through
sess.run(net.result, {image_input: x, keep_prob: 1}), decode + suppress on the GPU."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ssdutils   # noqa: E402
import synth      # noqa: E402
from ssdvgg import SSDVGG, Session   # noqa: E402


def main():
    ap = argparse.ArgumentParser(description='SSD inference')
    ap.add_argument('--name', default='test')
    ap.add_argument('--checkpoint', default='')
    ap.add_argument('--preset', default='vgg300')
    ap.add_argument('--threshold', type=float, default=0.5)
    ap.add_argument('--batch-size', type=int, default=32)
    ap.add_argument('--batches', type=int, default=1)
    ap.add_argument('--output-dir', default='')
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
        for b in range(args.batches):
            x = synth.images(b * args.batch_size, args.batch_size, preset.image_size.w)
            result = sess.run(net.result, feed_dict={net.image_input: x, net.keep_prob: 1})
            # infer.py:233-235: decode with no cap, suppress, keep the first 200 of the class-grouped list
            dets = [d[:200] for d in ssdutils.detect_batch(result, anchors, args.threshold, {}, None)]
            print('[i] batch %d: %s detections per image' % (b, [len(d) for d in dets][:8]))
            if args.output_dir:
                os.makedirs(args.output_dir, exist_ok=True)
                np.save(os.path.join(args.output_dir, 'result_%d.npy' % b), np.array(result))
    return 0


if __name__ == '__main__':
    sys.exit(main())
